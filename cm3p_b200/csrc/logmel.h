// Internal C++ interface of the log-mel front-end kernels (logmel.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cm3p {

int logmel_frames(const float* wave, const float* window, void* hi, void* lo, int batch, int64_t samples, int frames,
                  int n_fft, int hop, int ld, cudaStream_t s);
int logmel_power_mel(const float* spec, int64_t ld_spec, const float* filt, float* out, float* clip_max, int batch,
                     int frames, int bins, int mels, cudaStream_t s);
int logmel_finalize(float* out, const float* clip_max, int batch, int64_t per_clip, cudaStream_t s);

}  // namespace cm3p
