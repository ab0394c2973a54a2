// Internal C++ interface of the optimizer-step kernels (optim.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cm3p {

int muon_momentum(const float* g, float* buf, void* x, int64_t n, float momentum, int nesterov, float* sumsq,
                  cudaStream_t s);
int bf16_normalize(void* x, int64_t n, const float* sumsq, float eps, cudaStream_t s);
int bf16_axpy(void* out, int64_t ld_out, float a, const void* x, int64_t ld_x, const void* y, int64_t ld_y,
              int64_t rows, int64_t cols, cudaStream_t s);
int muon_apply(float* p, const void* x, int64_t n, float post_scale, float alpha, cudaStream_t s);
int adamw_step(float* p, const float* g, float* m1, float* m2, int64_t n, float beta1, float beta2, float eps,
               float decay, float step_size, cudaStream_t s);

}  // namespace cm3p
