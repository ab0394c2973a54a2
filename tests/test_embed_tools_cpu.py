"""The CPU restatement of the visualizer's analysis core (oracle/embed_tools_oracle.py) against the fixtures and
properties of the reference's own unit tests (visualizer/wasm/src/tests.rs; the Rust crate cannot be built here)."""
import numpy as np

from oracle import embed_tools_oracle as O

FIX = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [1.5, 2.5, 3.5], [10.0, 11.0, 12.0], [10.5, 11.5, 12.5]], dtype=np.float32)


def test_pca_fixtures():
    proj, mean, comps = O.calculate_pca(FIX)                      # tests.rs:19-25 shape
    assert proj.shape == (5, 2)
    assert O.calculate_pca(np.zeros((0, 0), dtype=np.float32))[0].shape == (0, 2)   # :28-31 empty input
    a, b = O.calculate_pca(FIX)[0], O.calculate_pca(FIX)[0]        # :34-43 deterministic
    assert np.array_equal(a, b)
    x = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [10, 10, 10]], dtype=np.float32)   # :201-222 outlier fixture
    p = O.calculate_pca(x)[0]
    assert np.linalg.norm(p[0] - p[3]) > np.linalg.norm(p[0] - p[1])
    rs = np.random.RandomState(1)
    _, _, comps = O.calculate_pca(rs.standard_normal((200, 16)).astype(np.float32) * np.linspace(3, 0.5, 16, dtype=np.float32))
    assert abs(float(comps[0] @ comps[1])) < 1e-4 and abs(float(np.linalg.norm(comps[1])) - 1) < 1e-4


def test_kmeans_fixtures():
    labels, _ = O.calculate_kmeans(FIX, 2, 42)                    # tests.rs:46-85
    assert labels.shape == (5,) and labels.dtype == np.int8
    assert set(labels.tolist()) <= {0, 1}
    assert labels[0] == labels[1] == labels[2] and labels[3] == labels[4] and labels[0] != labels[3]
    lab3, _ = O.calculate_kmeans(FIX, 3, 42)
    assert all(0 <= v < 3 for v in lab3.tolist())                 # :56-64 label range
    assert O.calculate_kmeans(np.zeros((0, 0), dtype=np.float32), 2, 42)[0].shape == (0,)   # :88-91


def test_normalize_fixtures():
    n = O.normalize_vectors(FIX)                                   # tests.rs:94-136
    assert n.shape == FIX.shape
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)
    z = O.normalize_vectors(np.array([[0, 0, 0], [1, 2, 3]], dtype=np.float32))
    assert np.array_equal(z[0], np.zeros(3, dtype=np.float32)) and abs(np.linalg.norm(z[1]) - 1) < 1e-5


def test_neighbor_fixtures():
    n = O.normalize_vectors(FIX)                                   # tests.rs:139-198
    idx, dist = O.find_nearest_neighbors(n, 0, 3)
    assert len(idx) == 3 and len(dist) == 3 and 0 not in idx.tolist()
    assert all(dist[i] <= dist[i + 1] for i in range(2))
    idx4, _ = O.find_nearest_neighbors(n, 0, 4)
    assert 1 in idx4[:2].tolist() or 2 in idx4[:2].tolist()
    e_idx, e_dist = O.find_nearest_neighbors(n, 999, 3)
    assert len(e_idx) == 0 and len(e_dist) == 0


def test_large_dataset_shapes():
    rs = np.random.RandomState(0)                                  # tests.rs:225-252
    x = rs.standard_normal((300, 64)).astype(np.float32)
    assert O.calculate_pca(x)[0].shape == (300, 2)
    assert O.calculate_kmeans(x, 5, 42)[0].shape == (300,)
    xn = O.normalize_vectors(x)
    assert O.find_nearest_neighbors(xn, 0, 10)[0].shape == (10,)
