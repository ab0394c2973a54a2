// Analysis kernels over the table of beatmap embeddings [n, d] fp32: row normalisation, brute-force cosine nearest
// neighbours, 2-component PCA by power iteration and k-means — what the reference's browser visualizer computes in
// its Rust -> WASM core (/root/reference/visualizer/wasm/src/lib.rs: normalize_vectors :371, find_nearest_neighbors
// :448, calculate_pca :82, calculate_kmeans :242) on the parquet written by extract_beatmap_embeddings.py.  On the GPU
// the table (244 K beatmaps x 512 = 500 MB) is streamed at HBM speed; every kernel here is bandwidth-bound, so rows
// are read with 16-byte loads by whole warps and reduced with shuffles.  All reductions have a fixed order (per-CTA
// partials summed by index), so results are bit-reproducible; ties are broken towards the lower index.
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "common.h"
#include "embed_tools.h"

namespace cm3p {
namespace {

constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------ normalise
// out[i] = x[i] / |x[i]|; all-zero rows stay zero (lib.rs:401-403)
__global__ void __launch_bounds__(THREADS)
normalize_rows_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* xr = x + row * d;
  float ss = 0.f;
  for (int k = lane; k < d; k += 32) ss += xr[k] * xr[k];
  ss = warp_sum(ss);
  const float inv = ss == 0.f ? 0.f : 1.f / sqrtf(ss);
  for (int k = lane; k < d; k += 32) out[row * d + k] = xr[k] * inv;
}

// ------------------------------------------------------------------------------------------------ column mean
// partial[b][k] = sum of column k over the rows of CTA b (rows strided by the grid); mean = sum_b partial / n
__global__ void __launch_bounds__(THREADS)
colsum_partial_kernel(const float* __restrict__ x, float* __restrict__ partial, int64_t n, int d) {
  extern __shared__ float sm[];  // [WARPS][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = lane; k < d; k += 32) sm[warp * d + k] = 0.f;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS + warp; row < n; row += static_cast<int64_t>(gridDim.x) * WARPS) {
    const float* xr = x + row * d;
    for (int k = lane; k < d; k += 32) sm[warp * d + k] += xr[k];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < d; k += THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += sm[w * d + k];
    partial[static_cast<int64_t>(blockIdx.x) * d + k] = s;
  }
}

// out[k] = scale * sum_b partial[b][k]  (fixed order)
__global__ void __launch_bounds__(THREADS)
reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out, int blocks, int d, float scale) {
  const int k = blockIdx.x * THREADS + threadIdx.x;
  if (k >= d) return;
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += partial[static_cast<int64_t>(b) * d + k];
  out[k] = s * scale;
}

// ------------------------------------------------------------------------------------------------ PCA power step
// One power iteration of the covariance operator without forming it (lib.rs:134-171):
//   score_i = <x_i - mean, ev>;   next = sum_i score_i (x_i - mean)
// Each row is read ONCE: the warp keeps it in registers between the dot product and the accumulation.
template <int MAXK>  // d <= 32 * MAXK
__global__ void __launch_bounds__(THREADS)
pca_power_partial_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ ev,
                         float* __restrict__ partial, int64_t n, int d) {
  extern __shared__ float sm[];  // [WARPS][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m[MAXK], e[MAXK], acc[MAXK];
#pragma unroll
  for (int j = 0; j < MAXK; ++j) {
    const int k = lane + 32 * j;
    m[j] = k < d ? mean[k] : 0.f;
    e[j] = k < d ? ev[k] : 0.f;
    acc[j] = 0.f;
  }
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS + warp; row < n; row += static_cast<int64_t>(gridDim.x) * WARPS) {
    const float* xr = x + row * d;
    float c[MAXK];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MAXK; ++j) {
      const int k = lane + 32 * j;
      c[j] = k < d ? xr[k] - m[j] : 0.f;
      s += c[j] * e[j];
    }
    s = warp_sum(s);
#pragma unroll
    for (int j = 0; j < MAXK; ++j) acc[j] += s * c[j];
  }
#pragma unroll
  for (int j = 0; j < MAXK; ++j) {
    const int k = lane + 32 * j;
    if (k < d) sm[warp * d + k] = acc[j];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < d; k += THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += sm[w * d + k];
    partial[static_cast<int64_t>(blockIdx.x) * d + k] = s;
  }
}

// ev <- next / |next| (kept if |next| == 0, lib.rs:163-169); with `ortho` also ev -= <u, ev> u first, then renormalise
// (the reference orthogonalises the second component once, AFTER its power iterations: lib.rs:173-186).  One CTA.
__global__ void __launch_bounds__(THREADS)
pca_finish_vector_kernel(const float* __restrict__ partial, int blocks, float* __restrict__ ev, const float* __restrict__ u,
                         int d, int mode) {
  // mode 0: ev = normalise(sum of partials) unless zero; mode 1: ev = normalise(ev - <u, ev> u) unless zero
  extern __shared__ float sm[];  // [d] + [THREADS]
  float* v = sm;
  float* red = sm + d;
  for (int k = threadIdx.x; k < d; k += THREADS) {
    float s;
    if (mode == 0) {
      s = 0.f;
      for (int b = 0; b < blocks; ++b) s += partial[static_cast<int64_t>(b) * d + k];
    } else {
      s = ev[k];
    }
    v[k] = s;
  }
  __syncthreads();
  auto block_dot = [&](const float* a, const float* b) -> float {
    float s = 0.f;
    for (int k = threadIdx.x; k < d; k += THREADS) s += a[k] * b[k];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = THREADS / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    const float r = red[0];
    __syncthreads();
    return r;
  };
  if (mode == 1) {
    const float dot = block_dot(u, v);
    for (int k = threadIdx.x; k < d; k += THREADS) v[k] -= dot * u[k];
    __syncthreads();
  }
  const float mag = sqrtf(block_dot(v, v));
  if (mag > 0.f) {
    const float inv = 1.f / mag;
    for (int k = threadIdx.x; k < d; k += THREADS) ev[k] = v[k] * inv;
  }
}

// proj[i] = (<x_i - mean, c0>, <x_i - mean, c1>)  (lib.rs:191-237)
__global__ void __launch_bounds__(THREADS)
pca_project_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ comp,
                   float* __restrict__ proj, int64_t n, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* xr = x + row * d;
  float a = 0.f, b = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float c = xr[k] - mean[k];
    a += c * comp[k];
    b += c * comp[d + k];
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) {
    proj[row * 2] = a;
    proj[row * 2 + 1] = b;
  }
}

// ------------------------------------------------------------------------------------------------ kNN
// dist[i] = 1 - <x_i, x_q> for normalised rows (lib.rs:466-474); dist[q] = +inf (the query itself is skipped)
__global__ void __launch_bounds__(THREADS)
cosine_distance_kernel(const float* __restrict__ x, float* __restrict__ dist, int64_t n, int d, int64_t q) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* xr = x + row * d;
  const float* xq = x + q * d;
  float s = 0.f;
  for (int k = lane; k < d; k += 32) s += xr[k] * xq[k];
  s = warp_sum(s);
  if (lane == 0) dist[row] = row == q ? FLT_MAX : 1.f - s;
}

// The `k` smallest (value, index) pairs of `vals[lo, hi)` in ascending (value, index) order, by k rounds of block
// argmin over a shared copy.  chunked: CTA b selects from its slice and writes candidates [b][k]; the final pass
// (one CTA) selects from the candidate list through `cand_idx`.
constexpr int SEL_CHUNK = 4096;
__global__ void __launch_bounds__(THREADS)
select_smallest_kernel(const float* __restrict__ vals, const int64_t* __restrict__ src_idx, int64_t count, int k,
                       float* __restrict__ out_val, int64_t* __restrict__ out_idx) {
  __shared__ float sv[SEL_CHUNK];
  __shared__ float rv[THREADS];
  __shared__ int ri[THREADS];
  const int64_t lo = static_cast<int64_t>(blockIdx.x) * SEL_CHUNK;
  const int64_t left = count - lo;
  const int m = static_cast<int>(left < SEL_CHUNK ? left : SEL_CHUNK);
  for (int i = threadIdx.x; i < SEL_CHUNK; i += THREADS) sv[i] = i < m ? vals[lo + i] : FLT_MAX;
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    float bv = FLT_MAX;
    int bi = SEL_CHUNK;
    for (int i = threadIdx.x; i < m; i += THREADS) {
      const float v = sv[i];
      if (v < bv) { bv = v; bi = i; }  // strided scan keeps the lowest index among equals per thread
    }
    rv[threadIdx.x] = bv;
    ri[threadIdx.x] = bi;
    __syncthreads();
    for (int o = THREADS / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        const float ov = rv[threadIdx.x + o];
        const int oi = ri[threadIdx.x + o];
        if (ov < rv[threadIdx.x] || (ov == rv[threadIdx.x] && oi < ri[threadIdx.x])) {
          rv[threadIdx.x] = ov;
          ri[threadIdx.x] = oi;
        }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const int bi0 = ri[0];
      const int64_t o = static_cast<int64_t>(blockIdx.x) * k + r;
      if (bi0 < m) {
        out_val[o] = rv[0];
        out_idx[o] = src_idx ? src_idx[lo + bi0] : lo + bi0;
        sv[bi0] = FLT_MAX;  // taken
      } else {
        out_val[o] = FLT_MAX;
        out_idx[o] = -1;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ k-means
// dist_min[i] = min(dist_min[i], |x_i - c|^2)   (k-means++-style farthest-point seeding, lib.rs:262-276)
__global__ void __launch_bounds__(THREADS)
kmeans_update_seed_distance_kernel(const float* __restrict__ x, const float* __restrict__ centroid,
                                   float* __restrict__ dist_min, int64_t n, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* xr = x + row * d;
  float s = 0.f;
  for (int k = lane; k < d; k += 32) {
    const float t = xr[k] - centroid[k];
    s += t * t;
  }
  s = warp_sum(s);
  if (lane == 0 && s < dist_min[row]) dist_min[row] = s;
}

// arg max with the reference's scan semantics: the first index whose value exceeds every earlier one (values <= 0
// never win: `max_dist` starts at 0, lib.rs:279-286).  partial per CTA, then one CTA over the partials.
__global__ void __launch_bounds__(THREADS)
argmax_first_kernel(const float* __restrict__ vals, const int64_t* __restrict__ src_idx, int64_t count,
                    float* __restrict__ out_val, int64_t* __restrict__ out_idx) {
  __shared__ float rv[THREADS];
  __shared__ int64_t ri[THREADS];
  const int64_t per = (count + gridDim.x - 1) / gridDim.x;
  const int64_t lo = per * blockIdx.x, hi = (lo + per < count) ? lo + per : count;
  float bv = 0.f;
  int64_t bi = -1;
  for (int64_t i = lo + threadIdx.x; i < hi; i += THREADS) {
    const float v = vals[i];
    const int64_t id = src_idx ? src_idx[i] : i;
    if (v > bv || (v == bv && bi >= 0 && id < bi)) { bv = v; bi = id; }
  }
  rv[threadIdx.x] = bv;
  ri[threadIdx.x] = bi;
  __syncthreads();
  for (int o = THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float ov = rv[threadIdx.x + o];
      const int64_t oi = ri[threadIdx.x + o];
      if (oi >= 0 && (ov > rv[threadIdx.x] || (ov == rv[threadIdx.x] && (ri[threadIdx.x] < 0 || oi < ri[threadIdx.x])))) {
        rv[threadIdx.x] = ov;
        ri[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out_val[blockIdx.x] = rv[0];
    out_idx[blockIdx.x] = ri[0];
  }
}

__global__ void copy_row_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, float* __restrict__ dst, int d) {
  const int64_t r = *idx < 0 ? 0 : *idx;  // no positive distance left: the reference keeps index 0
  for (int k = threadIdx.x; k < d; k += blockDim.x) dst[k] = x[r * d + k];
}

// label[i] = arg min_c |x_i - centroid_c|^2 (first minimum, lib.rs:300-318); changed += (label changed)
__global__ void __launch_bounds__(THREADS)
kmeans_assign_kernel(const float* __restrict__ x, const float* __restrict__ centroids, int8_t* __restrict__ labels,
                     int* __restrict__ changed, int64_t n, int d, int k) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * WARPS + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* xr = x + row * d;
  float best = FLT_MAX;
  int best_c = labels[row];
  for (int c = 0; c < k; ++c) {
    const float* cr = centroids + static_cast<int64_t>(c) * d;
    float s = 0.f;
    for (int j = lane; j < d; j += 32) {
      const float t = xr[j] - cr[j];
      s += t * t;
    }
    s = warp_sum(s);
    if (s < best) { best = s; best_c = c; }
  }
  if (lane == 0 && labels[row] != best_c) {
    labels[row] = static_cast<int8_t>(best_c);
    atomicAdd(changed, 1);  // a count: order-independent
  }
}

// Per-cluster sums with a fixed reduction order: CTA (c, b) sums the rows of cluster c among rows b, b+grid.y, ...
// (warps stride rows, lanes stride columns), partial [c][b][d] and count [c][b]; kmeans_update_kernel folds them.
__global__ void __launch_bounds__(THREADS)
kmeans_cluster_partial_kernel(const float* __restrict__ x, const int8_t* __restrict__ labels, float* __restrict__ partial,
                              int* __restrict__ pcount, int64_t n, int d) {
  extern __shared__ float sm[];  // [WARPS][d]
  __shared__ int cnt[WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x;
  for (int k = lane; k < d; k += 32) sm[warp * d + k] = 0.f;
  int mine = 0;
  for (int64_t row = static_cast<int64_t>(blockIdx.y) * WARPS + warp; row < n; row += static_cast<int64_t>(gridDim.y) * WARPS) {
    if (labels[row] != c) continue;  // warp-uniform
    ++mine;
    const float* xr = x + row * d;
    for (int k = lane; k < d; k += 32) sm[warp * d + k] += xr[k];
  }
  if (lane == 0) cnt[warp] = mine;
  __syncthreads();
  const int64_t slot = static_cast<int64_t>(c) * gridDim.y + blockIdx.y;
  for (int k = threadIdx.x; k < d; k += THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += sm[w * d + k];
    partial[slot * d + k] = s;
  }
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < WARPS; ++w) t += cnt[w];
    pcount[slot] = t;
  }
}

// centroid_c = sum / count for non-empty clusters (empty ones keep their centroid, lib.rs:352-361)
__global__ void __launch_bounds__(THREADS)
kmeans_update_kernel(const float* __restrict__ partial, const int* __restrict__ pcount, float* __restrict__ centroids,
                     int blocks, int d) {
  const int c = blockIdx.x;
  __shared__ int total;
  if (threadIdx.x == 0) {
    int t = 0;
    for (int b = 0; b < blocks; ++b) t += pcount[static_cast<int64_t>(c) * blocks + b];
    total = t;
  }
  __syncthreads();
  if (total == 0) return;
  const float inv = 1.f / static_cast<float>(total);
  for (int k = threadIdx.x; k < d; k += THREADS) {
    float s = 0.f;
    for (int b = 0; b < blocks; ++b) s += partial[(static_cast<int64_t>(c) * blocks + b) * d + k];
    centroids[static_cast<int64_t>(c) * d + k] = s * inv;
  }
}

inline int row_blocks(int64_t n) { return static_cast<int>((n + WARPS - 1) / WARPS); }
inline int stream_blocks(int64_t n) {
  const int sms = num_sms() > 0 ? num_sms() : 148;
  const int64_t want = (n + WARPS - 1) / WARPS;
  return static_cast<int>(want < 4 * sms ? (want < 1 ? 1 : want) : 4 * sms);
}

}  // namespace

int embed_normalize(const float* x, float* out, int64_t n, int d, cudaStream_t stream) {
  CM3P_REQUIRE(x && out && n > 0 && d > 0, kBadShape, "normalize_vectors: empty table");
  normalize_rows_kernel<<<row_blocks(n), THREADS, 0, stream>>>(x, out, n, d);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int embed_pca_workspace_floats(int64_t n, int d) { return stream_blocks(n) * d + 2 * d; }

int embed_pca2(const float* x, int64_t n, int d, const float* init, int iterations, float* mean, float* components,
               float* proj, float* ws, cudaStream_t stream) {
  CM3P_REQUIRE(x && init && mean && components && proj && ws && n > 0 && d > 0, kBadShape, "pca: empty table / null");
  CM3P_REQUIRE(d <= 1024, kBadShape, "pca: d=%d > 1024 unsupported", d);
  const int blocks = stream_blocks(n);
  float* partial = ws;
  const size_t row_smem = static_cast<size_t>(WARPS) * d * sizeof(float);
  CM3P_CUDA_TRY(cudaFuncSetAttribute(colsum_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  colsum_partial_kernel<<<blocks, THREADS, row_smem, stream>>>(x, partial, n, d);
  reduce_partials_kernel<<<(d + THREADS - 1) / THREADS, THREADS, 0, stream>>>(partial, mean, blocks, d,
                                                                              1.f / static_cast<float>(n));
  CM3P_CUDA_TRY(cudaGetLastError());
  CM3P_CUDA_TRY(cudaMemcpyAsync(components, init, 2 * static_cast<size_t>(d) * sizeof(float), cudaMemcpyDeviceToDevice,
                                stream));
  const size_t fin_smem = (static_cast<size_t>(d) + THREADS) * sizeof(float);
  for (int c = 0; c < 2; ++c) {
    float* ev = components + static_cast<size_t>(c) * d;
    // the random start vector is normalised first (lib.rs:126-131): mode 1 with u = ev itself would zero it, so
    // normalise through mode 0 over a single "partial" = the vector
    pca_finish_vector_kernel<<<1, THREADS, fin_smem, stream>>>(ev, 1, ev, nullptr, d, 0);
    for (int it = 0; it < iterations; ++it) {
      if (d <= 256) {
        CM3P_CUDA_TRY(cudaFuncSetAttribute(pca_power_partial_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        pca_power_partial_kernel<8><<<blocks, THREADS, row_smem, stream>>>(x, mean, ev, partial, n, d);
      } else if (d <= 512) {
        CM3P_CUDA_TRY(cudaFuncSetAttribute(pca_power_partial_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        pca_power_partial_kernel<16><<<blocks, THREADS, row_smem, stream>>>(x, mean, ev, partial, n, d);
      } else {
        CM3P_CUDA_TRY(cudaFuncSetAttribute(pca_power_partial_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        pca_power_partial_kernel<32><<<blocks, THREADS, row_smem, stream>>>(x, mean, ev, partial, n, d);
      }
      pca_finish_vector_kernel<<<1, THREADS, fin_smem, stream>>>(partial, blocks, ev, nullptr, d, 0);
    }
    if (c == 1) pca_finish_vector_kernel<<<1, THREADS, fin_smem, stream>>>(nullptr, 0, ev, components, d, 1);
    CM3P_CUDA_TRY(cudaGetLastError());
  }
  pca_project_kernel<<<row_blocks(n), THREADS, 0, stream>>>(x, mean, components, proj, n, d);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int64_t embed_knn_workspace_bytes(int64_t n, int k) {
  const int64_t chunks = (n + SEL_CHUNK - 1) / SEL_CHUNK;
  return n * 4 + chunks * k * (4 + 8) + 64;
}

int embed_knn(const float* xn, int64_t n, int d, int64_t query, int k, int64_t* out_idx, float* out_dist, void* ws,
              cudaStream_t stream) {
  CM3P_REQUIRE(xn && out_idx && out_dist && ws && n > 1 && d > 0, kBadShape, "knn: empty table / null");
  CM3P_REQUIRE(query >= 0 && query < n, kBadShape, "knn: query index %lld out of range", (long long)query);
  CM3P_REQUIRE(k > 0 && k <= n - 1 && k <= 1024, kBadShape, "knn: k=%d must be in [1, min(n-1, 1024)]", k);
  const int64_t chunks = (n + SEL_CHUNK - 1) / SEL_CHUNK;
  CM3P_REQUIRE(chunks * k <= SEL_CHUNK, kBadShape, "knn: n=%lld with k=%d needs a third selection level",
               (long long)n, k);
  float* dist = reinterpret_cast<float*>(ws);
  int64_t* cand_idx = reinterpret_cast<int64_t*>(reinterpret_cast<uint8_t*>(ws) + ((n * 4 + 15) / 16) * 16);
  float* cand_val = reinterpret_cast<float*>(cand_idx + chunks * k);
  cosine_distance_kernel<<<row_blocks(n), THREADS, 0, stream>>>(xn, dist, n, d, query);
  select_smallest_kernel<<<static_cast<unsigned>(chunks), THREADS, 0, stream>>>(dist, nullptr, n, k, cand_val, cand_idx);
  select_smallest_kernel<<<1, THREADS, 0, stream>>>(cand_val, cand_idx, chunks * k, k, out_dist, out_idx);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int64_t embed_kmeans_workspace_bytes(int64_t n, int d, int k) {
  const int blocks = 64;
  return n * 4 + static_cast<int64_t>(k) * blocks * d * 4 + static_cast<int64_t>(k) * blocks * 4 + 1024 * (4 + 8) + 256;
}

int embed_kmeans(const float* x, int64_t n, int d, int k, int64_t first_index, int iterations, float* centroids,
                 int8_t* labels, int* changed_per_iter, void* ws, cudaStream_t stream) {
  CM3P_REQUIRE(x && centroids && labels && changed_per_iter && ws && n > 0 && d > 0, kBadShape, "kmeans: empty / null");
  CM3P_REQUIRE(k > 0 && k <= 127, kBadShape, "kmeans: k=%d must be in [1, 127] (labels are int8 like the reference's)", k);
  CM3P_REQUIRE(first_index >= 0 && first_index < n && d <= 2048, kBadShape, "kmeans: bad first index or d > 2048");
  constexpr int blocks = 64;
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);
  float* dist_min = reinterpret_cast<float*>(w);
  w += ((n * 4 + 15) / 16) * 16;
  float* partial = reinterpret_cast<float*>(w);
  w += static_cast<int64_t>(k) * blocks * d * 4;
  int* pcount = reinterpret_cast<int*>(w);
  w += ((static_cast<int64_t>(k) * blocks * 4 + 15) / 16) * 16;
  int64_t* am_idx = reinterpret_cast<int64_t*>(w);
  float* am_val = reinterpret_cast<float*>(am_idx + 1024);
  // ---- seeding: first centroid given, then the point farthest from its nearest centroid so far (lib.rs:250-291)
  {
    // dist_min = +inf
    CM3P_CUDA_TRY(cudaMemsetAsync(dist_min, 0x7f, n * 4, stream));  // 0x7f7f7f7f ~ 3.4e38
    CM3P_CUDA_TRY(cudaMemcpyAsync(centroids, x + first_index * d, static_cast<size_t>(d) * 4, cudaMemcpyDeviceToDevice, stream));
    const int ab = static_cast<int>(n < 1000 * 256 ? (n + 255) / 256 : 1000);
    for (int i = 1; i < k; ++i) {
      kmeans_update_seed_distance_kernel<<<row_blocks(n), THREADS, 0, stream>>>(x, centroids + static_cast<int64_t>(i - 1) * d,
                                                                              dist_min, n, d);
      argmax_first_kernel<<<ab, THREADS, 0, stream>>>(dist_min, nullptr, n, am_val, am_idx);
      argmax_first_kernel<<<1, THREADS, 0, stream>>>(am_val, am_idx, ab, am_val + 1024 - 8, am_idx + 1024 - 8);
      copy_row_kernel<<<1, 256, 0, stream>>>(x, am_idx + 1024 - 8, centroids + static_cast<int64_t>(i) * d, d);
    }
    CM3P_CUDA_TRY(cudaGetLastError());
  }
  // ---- Lloyd iterations (lib.rs:295-363); the caller reads changed_per_iter to apply the reference's early stop
  CM3P_CUDA_TRY(cudaMemsetAsync(labels, 0, n, stream));
  CM3P_CUDA_TRY(cudaMemsetAsync(changed_per_iter, 0, static_cast<size_t>(iterations) * 4, stream));
  const size_t row_smem = static_cast<size_t>(WARPS) * d * sizeof(float);
  CM3P_CUDA_TRY(cudaFuncSetAttribute(kmeans_cluster_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  for (int it = 0; it < iterations; ++it) {
    kmeans_assign_kernel<<<row_blocks(n), THREADS, 0, stream>>>(x, centroids, labels, changed_per_iter + it, n, d, k);
    kmeans_cluster_partial_kernel<<<dim3(k, blocks), THREADS, row_smem, stream>>>(x, labels, partial, pcount, n, d);
    kmeans_update_kernel<<<k, THREADS, 0, stream>>>(partial, pcount, centroids, blocks, d);
  }
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

}  // namespace cm3p
