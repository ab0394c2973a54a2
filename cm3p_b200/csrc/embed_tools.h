// Internal C++ interface of the embedding-table analysis kernels (embed_tools.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cm3p {

int embed_normalize(const float* x, float* out, int64_t n, int d, cudaStream_t stream);
int embed_pca_workspace_floats(int64_t n, int d);
int embed_pca2(const float* x, int64_t n, int d, const float* init, int iterations, float* mean, float* components,
               float* proj, float* ws, cudaStream_t stream);
int64_t embed_knn_workspace_bytes(int64_t n, int k);
int embed_knn(const float* xn, int64_t n, int d, int64_t query, int k, int64_t* out_idx, float* out_dist, void* ws,
              cudaStream_t stream);
int64_t embed_kmeans_workspace_bytes(int64_t n, int d, int k);
int embed_kmeans(const float* x, int64_t n, int d, int k, int64_t first_index, int iterations, float* centroids,
                 int8_t* labels, int* changed_per_iter, void* ws, cudaStream_t stream);

}  // namespace cm3p
