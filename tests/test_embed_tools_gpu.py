"""Embedding-table analysis kernels on a B200 (through the C ABI) against the CPU restatement of the visualizer's
WASM core (oracle/embed_tools_oracle.py): the reference's own 5x3 fixtures, and seeded tables at production width."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import embed_tools_oracle as O

FIX = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [1.5, 2.5, 3.5], [10.0, 11.0, 12.0], [10.5, 11.5, 12.5]], dtype=np.float32)


def _table(n, d, seed=0, clusters=7):
    rs = np.random.RandomState(seed)
    centers = rs.standard_normal((clusters, d)).astype(np.float32) * 3
    x = centers[rs.randint(0, clusters, n)] + rs.standard_normal((n, d)).astype(np.float32)
    return (x * np.linspace(2.0, 0.5, d, dtype=np.float32)).astype(np.float32)


@pytest.mark.parametrize("n,d", [(5, 3), (1000, 64), (9000, 512)])
def test_normalize_and_neighbors(n, d):
    from cm3p_b200 import embedding_tools as E
    x = FIX if n == 5 else _table(n, d)
    if n > 5:
        x[3] = 0.0  # a zero row stays zero
    xn = E.normalize_vectors(torch.from_numpy(x).cuda())
    want = O.normalize_vectors(x)
    torch.testing.assert_close(xn.cpu(), torch.from_numpy(want), rtol=1e-5, atol=1e-6)
    for q, k in ((0, 3), (n - 1, min(10, n - 1))):
        idx, dist = E.find_nearest_neighbors(xn, q, k)
        w_idx, w_dist = O.find_nearest_neighbors(want, q, k)
        assert len(idx) == k and q not in idx.tolist()
        assert bool((dist[1:] >= dist[:-1]).all())
        torch.testing.assert_close(dist.cpu(), torch.from_numpy(w_dist), rtol=0, atol=2e-6)
        # same neighbours wherever the distances are separated by more than rounding
        gaps = np.abs(np.diff(np.concatenate([w_dist, [np.inf]])))
        for j in range(k):
            if gaps[j] > 1e-5 and (j == 0 or gaps[j - 1] > 1e-5):
                assert int(idx[j]) == int(w_idx[j]), (q, j)
    e_idx, e_dist = E.find_nearest_neighbors(xn, n + 5, 3)
    assert e_idx.numel() == 0 and e_dist.numel() == 0


@pytest.mark.parametrize("n,d", [(4, 3), (2000, 64), (6000, 512)])
def test_pca(n, d):
    from cm3p_b200 import embedding_tools as E
    x = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [10, 10, 10]], dtype=np.float32) if n == 4 else _table(n, d, seed=1)
    proj, mean, comps = E.calculate_pca(torch.from_numpy(x).cuda(), return_basis=True)
    w_proj, w_mean, w_comps = O.calculate_pca(x)
    assert proj.shape == (n, 2)
    torch.testing.assert_close(mean.cpu(), torch.from_numpy(w_mean), rtol=1e-4, atol=1e-5)
    # components up to the summation order of fp32 reductions; sign is fixed by the shared start vectors
    c, wc = comps.cpu().numpy(), w_comps
    assert abs(float(c[0] @ wc[0])) > 0.999
    if n != 4:  # (in the 4-point fixture the un-deflated second vector collapses onto the first: nothing left to orthogonalise)
        assert abs(float(c[0] @ c[1])) < 1e-3 and abs(float(np.linalg.norm(c[1])) - 1) < 1e-4
    p = proj.cpu().numpy()
    scale = np.abs(w_proj[:, 0]).max()
    assert np.abs(p[:, 0] - w_proj[:, 0]).max() <= 2e-3 * scale
    if n == 4:  # the reference's outlier fixture (tests.rs:201-222)
        assert np.linalg.norm(p[0] - p[3]) > np.linalg.norm(p[0] - p[1])
    else:
        assert abs(float(c[1] @ wc[1])) > 0.99
        assert np.abs(p[:, 1] - w_proj[:, 1]).max() <= 2e-2 * np.abs(w_proj[:, 1]).max()
    # bit-reproducible
    proj2 = E.calculate_pca(torch.from_numpy(x).cuda())
    assert torch.equal(proj, proj2)


@pytest.mark.parametrize("n,d,k", [(5, 3, 2), (3000, 64, 7), (5000, 512, 20)])
def test_kmeans(n, d, k):
    from cm3p_b200 import embedding_tools as E
    x = FIX if n == 5 else _table(n, d, seed=2, clusters=k)
    labels, cent = E.calculate_kmeans(torch.from_numpy(x).cuda(), k, seed=42, return_centroids=True)
    w_labels, w_cent = O.calculate_kmeans(x, k, 42)
    assert labels.dtype == torch.int8 and labels.shape == (n,)
    lab = labels.cpu().numpy()
    assert lab.min() >= 0 and lab.max() < k
    if n == 5:  # tests.rs:68-85
        assert lab[0] == lab[1] == lab[2] and lab[3] == lab[4] and lab[0] != lab[3]
    # same partition as the oracle (identical seeding rule; assignments only differ where two centroids are
    # equidistant to rounding)
    agree = float((lab == w_labels).mean())
    assert agree >= 0.995, agree
    torch.testing.assert_close(cent.cpu(), torch.from_numpy(w_cent), rtol=1e-3, atol=1e-3)
    labels2 = E.calculate_kmeans(torch.from_numpy(x).cuda(), k, seed=42)
    assert torch.equal(labels, labels2)
