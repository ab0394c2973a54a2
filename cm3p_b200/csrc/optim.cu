// Optimizer-step kernels for the reference's Muon optimizer (utils/muon_utils.py:35-57 Newton-Schulz,
// :138-203 step): everything that is not a GEMM.  The three GEMMs per Newton-Schulz iteration run on
// gemm_bf16_sm100_kernel.  bf16 roundings follow the reference's `X = G.bfloat16()` arithmetic.
#include <cuda_bf16.h>
#include <math.h>

#include "common.h"
#include "optim.h"
#include "ptx.cuh"

namespace cm3p {
namespace {

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int k = 0; k < 8; ++k) t += red[k];
  return t;  // valid in thread 0
}

// buf = buf * momentum + g;  u = nesterov ? g + momentum * buf : buf;  x = bf16(u);  sumsq += sum x^2
// (muon_utils.py:152-156, :47)
__global__ void __launch_bounds__(256)
muon_momentum_kernel(const float* __restrict__ g, float* __restrict__ buf, __nv_bfloat16* __restrict__ x, int64_t n,
                     float momentum, int nesterov, float* __restrict__ sumsq) {
  float acc = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i];
    const float b = buf[i] * momentum + gi;
    buf[i] = b;
    const float u = nesterov ? gi + momentum * b : b;
    const __nv_bfloat16 xb = __float2bfloat16(u);
    x[i] = xb;
    const float xf = __bfloat162float(xb);
    acc += xf * xf;
  }
  const float t = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(sumsq, t);
}

// X /= (X.norm() + eps) with the reference's bf16 scalar arithmetic (muon_utils.py:48)
__global__ void __launch_bounds__(256)
bf16_normalize_kernel(__nv_bfloat16* __restrict__ x, int64_t n, const float* __restrict__ sumsq, float eps) {
  const float den = bf16_round(bf16_round(sqrtf(*sumsq)) + eps);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    x[i] = __float2bfloat16(__bfloat162float(x[i]) / den);
}

// out = bf16( bf16(a * x) + y )   (y may be null: out = bf16(a * x)).  Rows of `cols` elements with pitches.
__global__ void __launch_bounds__(256)
bf16_axpy_kernel(__nv_bfloat16* __restrict__ out, int64_t ld_out, float a, const __nv_bfloat16* __restrict__ x,
                 int64_t ld_x, const __nv_bfloat16* __restrict__ y, int64_t ld_y, int64_t rows, int64_t cols) {
  const int64_t n = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols, c = i % cols;
    float v = bf16_round(a * __bfloat162float(x[r * ld_x + c]));
    if (y) v += __bfloat162float(y[r * ld_y + c]);
    out[r * ld_out + c] = __float2bfloat16(v);
  }
}

// p += alpha * float( bf16(x * post_scale) )   (muon_utils.py:164-167)
__global__ void __launch_bounds__(256)
muon_apply_kernel(float* __restrict__ p, const __nv_bfloat16* __restrict__ x, int64_t n, float post_scale,
                  float alpha) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] += alpha * bf16_round(__bfloat162float(x[i]) * post_scale);
}

// The reference's internal AdamW (muon_utils.py:179-203), quirks included:
//   m1 = lerp(m1, g, 1-b1); m2 = lerp(m2, g^2, 1-b2); u = m1 / (eps + sqrt(m2));
//   p = p * decay - step_size * u        (decay = 1 - adamw_lr*wd, step_size = lr / scale)
__global__ void __launch_bounds__(256)
adamw_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m1, float* __restrict__ m2,
                  int64_t n, float beta1, float beta2, float eps, float decay, float step_size) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i];
    const float a = m1[i] + (gi - m1[i]) * (1.f - beta1);
    const float b = m2[i] + (gi * gi - m2[i]) * (1.f - beta2);
    m1[i] = a;
    m2[i] = b;
    p[i] = p[i] * decay - step_size * (a / (eps + sqrtf(b)));
  }
}

inline int grid_for(int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms() > 0 ? num_sms() : 148) * 8;
  if (g > cap) g = cap;
  return static_cast<int>(g < 1 ? 1 : g);
}

}  // namespace

int muon_momentum(const float* g, float* buf, void* x, int64_t n, float momentum, int nesterov, float* sumsq,
                  cudaStream_t s) {
  if (n == 0) return kOk;
  muon_momentum_kernel<<<grid_for(n), 256, 0, s>>>(g, buf, reinterpret_cast<__nv_bfloat16*>(x), n, momentum, nesterov,
                                                   sumsq);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}
int bf16_normalize(void* x, int64_t n, const float* sumsq, float eps, cudaStream_t s) {
  if (n == 0) return kOk;
  bf16_normalize_kernel<<<grid_for(n), 256, 0, s>>>(reinterpret_cast<__nv_bfloat16*>(x), n, sumsq, eps);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}
int bf16_axpy(void* out, int64_t ld_out, float a, const void* x, int64_t ld_x, const void* y, int64_t ld_y,
              int64_t rows, int64_t cols, cudaStream_t s) {
  if (rows * cols == 0) return kOk;
  bf16_axpy_kernel<<<grid_for(rows * cols), 256, 0, s>>>(reinterpret_cast<__nv_bfloat16*>(out), ld_out, a,
                                                         reinterpret_cast<const __nv_bfloat16*>(x), ld_x,
                                                         reinterpret_cast<const __nv_bfloat16*>(y), ld_y, rows, cols);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}
int muon_apply(float* p, const void* x, int64_t n, float post_scale, float alpha, cudaStream_t s) {
  if (n == 0) return kOk;
  muon_apply_kernel<<<grid_for(n), 256, 0, s>>>(p, reinterpret_cast<const __nv_bfloat16*>(x), n, post_scale, alpha);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}
int adamw_step(float* p, const float* g, float* m1, float* m2, int64_t n, float beta1, float beta2, float eps,
               float decay, float step_size, cudaStream_t s) {
  if (n == 0) return kOk;
  adamw_step_kernel<<<grid_for(n), 256, 0, s>>>(p, g, m1, m2, n, beta1, beta2, eps, decay, step_size);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

}  // namespace cm3p
