"""CPU restatement of the visualizer's analysis core.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/visualizer/wasm/src/lib.rs function by function, in numpy fp32:
  normalize_vectors       :371-432   (zero rows stay zero)
  find_nearest_neighbors  :448-488   (distance = 1 - dot, the query excluded, ascending; ties by index here)
  calculate_pca           :82-237    (mean, 2 x (random unit start, 8 power iterations of X_c^T X_c WITHOUT deflation,
                                      second vector orthogonalised once at the end), projection)
  calculate_kmeans        :242-365   (first centroid = LCG(seed) % n, farthest-point seeding, <= 10 Lloyd iterations,
                                      first-minimum assignment, empty clusters keep their centroid)
  simple_random           :7-10      (the native-build LCG behind the PCA start vectors; the WASM build uses Math.random)

Parity status: the reference is Rust -> wasm32 and there is no cargo / rustc in this image, so it cannot be run here.
The restatement is pinned instead to every fixture and property of the reference's own unit tests
(visualizer/wasm/src/tests.rs: shapes, label range, the 5x3 clustering / neighbour fixtures, unit length, zero vector,
self-exclusion, sortedness, the PCA outlier fixture) in tests/test_embed_tools_cpu.py.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def simple_random(state: int):
    state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
    return state, F(F(state) / F(0xFFFFFFFF))


def pca_start_vectors(d: int, state: int = 12345) -> np.ndarray:
    """The two un-normalised start vectors of the native build (lib.rs:112-124): one LCG stream, `- 0.5`."""
    out = np.zeros((2, d), dtype=F)
    for c in range(2):
        for j in range(d):
            state, r = simple_random(state)
            out[c, j] = r - F(0.5)
    return out


def normalize_vectors(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=F)
    ss = (x * x).sum(axis=1, dtype=F)
    inv = np.where(ss == 0, F(0), F(1) / np.sqrt(np.where(ss == 0, F(1), ss), dtype=F)).astype(F)
    return (x * inv[:, None]).astype(F)


def find_nearest_neighbors(xn: np.ndarray, query: int, k: int):
    n = xn.shape[0]
    if query >= n:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=F)
    dist = (F(1) - xn @ xn[query]).astype(F)
    idx = np.array([i for i in range(n) if i != query], dtype=np.int64)
    order = np.lexsort((idx, dist[idx]))[:min(k, n - 1)]
    return idx[order], dist[idx][order]


def calculate_pca(x: np.ndarray, init: np.ndarray | None = None, iterations: int = 8):
    x = np.asarray(x, dtype=F)
    n, d = x.shape
    if n == 0 or d == 0:
        return np.zeros((0, 2), dtype=F), None, None
    mean = (x.sum(axis=0, dtype=F) * F(1.0 / n)).astype(F)
    xc = (x - mean).astype(F)
    init = pca_start_vectors(d) if init is None else np.asarray(init, dtype=F)
    comps = []
    for c in range(2):
        ev = init[c].copy()
        ev = (ev / np.sqrt((ev * ev).sum(dtype=F))).astype(F)
        for _ in range(iterations):
            score = (xc @ ev).astype(F)
            nxt = (xc.T @ score).astype(F)
            mag = np.sqrt((nxt * nxt).sum(dtype=F))
            if mag > 0:
                ev = (nxt / mag).astype(F)
        if c == 1:
            u = comps[0]
            ev = (ev - (u @ ev) * u).astype(F)
            mag = np.sqrt((ev * ev).sum(dtype=F))
            if mag > 0:
                ev = (ev / mag).astype(F)
        comps.append(ev)
    comps = np.stack(comps)
    return (xc @ comps.T).astype(F), mean, comps


def kmeans_first_index(seed: int, n: int) -> int:
    return ((seed * 1664525 + 1013904223) & 0xFFFFFFFF) % n


def calculate_kmeans(x: np.ndarray, k: int, seed: int, iterations: int = 10):
    x = np.asarray(x, dtype=F)
    n, d = x.shape
    if n == 0 or k == 0:
        return np.zeros(0, dtype=np.int8), None
    cent = np.zeros((k, d), dtype=F)
    cent[0] = x[kmeans_first_index(seed, n)]
    dist = np.full(n, np.inf, dtype=F)
    for i in range(1, k):
        dd = ((x - cent[i - 1]) ** 2).sum(axis=1, dtype=F)
        dist = np.minimum(dist, dd)
        best, best_d = 0, F(0)
        for j in range(n):  # first index attaining the maximum (values <= 0 never win)
            if dist[j] > best_d:
                best_d, best = dist[j], j
        cent[i] = x[best]
    labels = np.zeros(n, dtype=np.int8)
    for it in range(iterations):
        d2 = ((x[:, None, :] - cent[None, :, :]) ** 2).sum(axis=2, dtype=F)  # [n, k]
        new = labels.copy()
        for i in range(n):
            best_c, best_d = int(labels[i]), np.inf
            for c in range(k):
                if d2[i, c] < best_d:
                    best_d, best_c = d2[i, c], c
            new[i] = best_c
        changed = int((new != labels).sum())
        labels = new
        if it > 0 and changed == 0:
            break
        for c in range(k):
            m = labels == c
            if m.any():
                cent[c] = (x[m].sum(axis=0, dtype=F) * F(1.0 / m.sum())).astype(F)
    return labels, cent
