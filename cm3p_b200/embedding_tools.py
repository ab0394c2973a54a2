"""Analysis of the embedding table on the GPU: row normalisation, cosine nearest neighbours, 2-component PCA and
k-means — the operations of the reference's browser visualizer (`visualizer/wasm/src/lib.rs`: normalize_vectors
:371, find_nearest_neighbors :448, calculate_pca :82, calculate_kmeans :242), which run there in Rust -> WASM over the
parquet `extract_beatmap_embeddings.py` writes.  Same function names, argument meaning and results (labels int8,
neighbours ascending by `1 - dot` with the query excluded, PCA by 8 un-deflated power iterations with the second
component orthogonalised at the end); the table stays in HBM and every pass is one streaming CUDA kernel through the
C ABI (include/cm3p_b200.h).  No CPU path.
"""
from __future__ import annotations

import torch

from . import _lib

_L32 = 0xFFFFFFFF


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _table(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("cm3p_b200.embedding_tools: the embedding table must be a CUDA tensor (no CPU path)")
    if x.dim() != 2:
        raise ValueError("embedding table must be [n_samples, n_features]")
    return x.float().contiguous()


def normalize_vectors(embeddings: torch.Tensor) -> torch.Tensor:
    x = _table(embeddings)
    out = torch.empty_like(x)
    if x.numel():
        _lib.check(_lib.load().cm3p_normalize_vectors(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], _stream()),
                   "cm3p_normalize_vectors")
    return out


def find_nearest_neighbors(normalized_embeddings: torch.Tensor, query_idx: int, n_neighbors: int):
    """-> (indices int64 [k], distances fp32 [k]); empty for an out-of-range query like the reference."""
    x = _table(normalized_embeddings)
    n, d = x.shape
    if query_idx >= n or query_idx < 0 or n < 2:
        return (torch.empty(0, dtype=torch.int64, device=x.device), torch.empty(0, dtype=torch.float32, device=x.device))
    k = min(int(n_neighbors), n - 1)
    lib = _lib.load()
    ws = torch.empty(int(lib.cm3p_knn_workspace_bytes(n, k)), dtype=torch.uint8, device=x.device)
    idx = torch.empty(k, dtype=torch.int64, device=x.device)
    dist = torch.empty(k, dtype=torch.float32, device=x.device)
    _lib.check(lib.cm3p_knn_cosine(x.data_ptr(), n, d, int(query_idx), k, idx.data_ptr(), dist.data_ptr(), ws.data_ptr(),
                                   ws.numel(), _stream()), "cm3p_knn_cosine")
    return idx, dist


def pca_start_vectors(n_features: int, state: int = 12345) -> torch.Tensor:
    """The start vectors of the reference's native build: one LCG stream, `simple_random() - 0.5` (lib.rs:7-10,112-124).
    (Its WASM build draws them from Math.random: any non-degenerate start converges to the same plane.)"""
    import numpy as np
    out = np.zeros((2, n_features), dtype=np.float32)
    for c in range(2):
        for j in range(n_features):
            state = (state * 1664525 + 1013904223) & _L32
            out[c, j] = np.float32(np.float32(state) / np.float32(_L32)) - np.float32(0.5)
    return torch.from_numpy(out)


def calculate_pca(embeddings: torch.Tensor, init: torch.Tensor | None = None, iterations: int = 8,
                  return_basis: bool = False):
    """-> projected [n, 2] fp32 (and (mean [d], components [2, d]) with return_basis)."""
    x = _table(embeddings)
    n, d = x.shape
    if n == 0 or d == 0:
        return torch.empty((0, 2), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    init = (pca_start_vectors(d) if init is None else init).to(device=x.device, dtype=torch.float32).contiguous()
    mean = torch.empty(d, dtype=torch.float32, device=x.device)
    comps = torch.empty((2, d), dtype=torch.float32, device=x.device)
    proj = torch.empty((n, 2), dtype=torch.float32, device=x.device)
    ws = torch.empty(int(lib.cm3p_pca2_workspace_floats(n, d)), dtype=torch.float32, device=x.device)
    _lib.check(lib.cm3p_pca2(x.data_ptr(), n, d, init.data_ptr(), int(iterations), mean.data_ptr(), comps.data_ptr(),
                             proj.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "cm3p_pca2")
    return (proj, mean, comps) if return_basis else proj


def calculate_kmeans(embeddings: torch.Tensor, k: int, seed: int = 42, iterations: int = 10,
                     return_centroids: bool = False):
    """-> labels int8 [n] (and centroids [k, d] with return_centroids)."""
    x = _table(embeddings)
    n, d = x.shape
    if n == 0 or k == 0:
        return torch.empty(0, dtype=torch.int8, device=x.device)
    lib = _lib.load()
    first = ((int(seed) * 1664525 + 1013904223) & _L32) % n  # lib.rs:254-255
    cent = torch.empty((k, d), dtype=torch.float32, device=x.device)
    labels = torch.empty(n, dtype=torch.int8, device=x.device)
    changed = torch.empty(iterations, dtype=torch.int32, device=x.device)
    ws = torch.empty(int(lib.cm3p_kmeans_workspace_bytes(n, d, k)), dtype=torch.uint8, device=x.device)
    _lib.check(lib.cm3p_kmeans(x.data_ptr(), n, d, int(k), int(first), int(iterations), cent.data_ptr(),
                               labels.data_ptr(), changed.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
               "cm3p_kmeans")
    return (labels, cent) if return_centroids else labels
