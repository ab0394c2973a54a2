// GPU log-mel front-end (what `WhisperFeatureExtractor` computes on the CPU for the reference's processor,
// cm3p/processing_cm3p.py:284-304; configs/train/default.yaml processor.audio_feature_extractor: n_fft 400,
// hop 160, 80 mel bins, 16 kHz): centre-padded (reflect) Hann frames -> DFT -> |.|^2 -> mel filter bank ->
// log10(max(., 1e-10)) -> max(., clip_max - 8) -> (x + 4) / 4.
// The DFT is a GEMM on the tensor cores: frames [B*F, 400] x [cos | -sin]^T, with both operands split into
// bf16 hi + lo parts (three tcgen05 GEMMs, fp32 accumulation) so the spectrum keeps ~fp32 accuracy.
#include <cuda_bf16.h>
#include <math.h>

#include "common.h"
#include "logmel.h"

namespace cm3p {
namespace {

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16(x);
  lo = __float2bfloat16(x - __bfloat162float(hi));
}

// frames[(b, f), n] = w[n] * x_b[reflect(f*hop + n - n_fft/2)], split into hi / lo bf16 matrices (pitch ld)
__global__ void __launch_bounds__(256)
frame_window_kernel(const float* __restrict__ wave, const float* __restrict__ window, __nv_bfloat16* __restrict__ hi,
                    __nv_bfloat16* __restrict__ lo, int64_t samples, int frames, int n_fft, int hop, int ld) {
  const int64_t row = blockIdx.x;  // b * frames + f
  const int b = static_cast<int>(row / frames), f = static_cast<int>(row % frames);
  const float* x = wave + static_cast<int64_t>(b) * samples;
  for (int n = threadIdx.x; n < ld; n += blockDim.x) {
    float v = 0.f;
    if (n < n_fft) {
      int64_t i = static_cast<int64_t>(f) * hop + n - n_fft / 2;
      if (i < 0) i = -i;                                  // numpy "reflect" (edge sample not repeated)
      if (i >= samples) i = 2 * (samples - 1) - i;
      v = x[i] * window[n];
    }
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[row * ld + n] = h;
    lo[row * ld + n] = l;
  }
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  // order-preserving integer view of IEEE floats
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// power spectrum -> mel -> log10; out[b, m, f] (frame contiguous), clip_max[b] = max over the clip
__global__ void __launch_bounds__(256)
power_mel_log_kernel(const float* __restrict__ spec, int64_t ld_spec, const float* __restrict__ filt,
                     float* __restrict__ out, float* __restrict__ clip_max, int frames, int bins, int mels) {
  extern __shared__ float sm[];  // filt [bins][mels] then power [32][bins + 1]
  float* s_filt = sm;
  float* s_pow = sm + bins * mels;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * 32;
  for (int i = threadIdx.x; i < bins * mels; i += blockDim.x) s_filt[i] = filt[i];
  for (int i = threadIdx.x; i < 32 * bins; i += blockDim.x) {
    const int fr = i / bins, k = i % bins;
    float p = 0.f;
    if (f0 + fr < frames) {
      const float* row = spec + (static_cast<int64_t>(b) * frames + f0 + fr) * ld_spec;
      const float re = row[k], im = row[bins + k];
      p = re * re + im * im;
    }
    s_pow[fr * (bins + 1) + k] = p;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < 32 * mels; i += blockDim.x) {
    const int m = i / 32, fr = i % 32;  // consecutive threads -> consecutive frames (coalesced store)
    if (f0 + fr >= frames) continue;
    float acc = 0.f;
    for (int k = 0; k < bins; ++k) acc += s_filt[k * mels + m] * s_pow[fr * (bins + 1) + k];
    const float v = log10f(fmaxf(acc, 1e-10f));
    out[(static_cast<int64_t>(b) * mels + m) * frames + f0 + fr] = v;
    mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > -INFINITY) atomic_max_float(clip_max + b, mx);
}

__global__ void __launch_bounds__(256)
logmel_finalize_kernel(float* __restrict__ out, const float* __restrict__ clip_max, int64_t per_clip, int64_t total) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float floor_v = clip_max[i / per_clip] - 8.0f;
    out[i] = (fmaxf(out[i], floor_v) + 4.0f) * 0.25f;
  }
}

}  // namespace

int logmel_frames(const float* wave, const float* window, void* hi, void* lo, int batch, int64_t samples, int frames,
                  int n_fft, int hop, int ld, cudaStream_t s) {
  CM3P_REQUIRE(ld >= n_fft && ld % 8 == 0 && samples > n_fft / 2, kBadShape, "logmel_frames: bad shape");
  frame_window_kernel<<<static_cast<unsigned>(static_cast<int64_t>(batch) * frames), 256, 0, s>>>(
      wave, window, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), samples, frames, n_fft,
      hop, ld);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int logmel_power_mel(const float* spec, int64_t ld_spec, const float* filt, float* out, float* clip_max, int batch,
                     int frames, int bins, int mels, cudaStream_t s) {
  const size_t smem = (static_cast<size_t>(bins) * mels + 32 * (bins + 1)) * sizeof(float);
  CM3P_ENSURE_DYN_SMEM(power_mel_log_kernel, 160 * 1024);
  CM3P_REQUIRE(smem <= 160 * 1024, kBadShape, "logmel: filter bank too large for shared memory");
  dim3 grid((frames + 31) / 32, batch);
  power_mel_log_kernel<<<grid, 256, smem, s>>>(spec, ld_spec, filt, out, clip_max, frames, bins, mels);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int logmel_finalize(float* out, const float* clip_max, int batch, int64_t per_clip, cudaStream_t s) {
  const int64_t total = per_clip * batch;
  int64_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  logmel_finalize_kernel<<<static_cast<unsigned>(g), 256, 0, s>>>(out, clip_max, per_clip, total);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

}  // namespace cm3p
