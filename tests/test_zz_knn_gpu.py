"""-m gpu: gs_knn3_mean_dist2 (the default distCUDA2 for CUDA tensors) against a numpy brute force (the definition: exact
3 nearest OTHER points, duplicates count at distance 0), and on an off-origin dense cloud against float64."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def brute(p):
    p = p.astype(np.float32)
    d = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1).astype(np.float32)
    np.fill_diagonal(d, np.inf)
    k = min(3, p.shape[0] - 1)
    return np.sort(d, axis=1)[:, :k].mean(axis=1).astype(np.float32) if k > 0 else np.zeros(p.shape[0], np.float32)


def brute64(p):
    """float64 distances of the float32 coordinates (the exact answer the fp32 paths are compared with)."""
    q = p.astype(np.float64)
    d = ((q[:, None, :] - q[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d, np.inf)
    return np.sort(d, axis=1)[:, :3].mean(axis=1)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 255, 256, 257, 3001])
def test_knn3_mean_dist2(n):
    from simple_knn import _C as knn
    rng = np.random.default_rng(n)
    p = rng.normal(size=(n, 3)).astype(np.float32)
    if n > 10:
        p[5] = p[7]                    # duplicates: distance 0 counts as a neighbour
        p[9] = p[7]
    t = torch.from_numpy(p).cuda()
    got = knn._dist2_kernel(t).cpu().numpy()
    np.testing.assert_allclose(got, brute(p), rtol=2e-6, atol=1e-7)
    assert np.array_equal(knn.distCUDA2(t).cpu().numpy(), got)          # the kernel IS the default for CUDA tensors
    np.testing.assert_allclose(got, knn._dist2_exact_torch(t).cpu().numpy(), rtol=2e-6, atol=1e-7)


def test_knn3_off_origin_dense_cloud():
    """3000 points at ~2.6e-3 spacing around (30,-20,15) -- the matmul form of cdist returns 0 for all."""
    from simple_knn import _C as knn
    rng = np.random.default_rng(7)
    p = (np.array([30.0, -20.0, 15.0]) + rng.uniform(-0.02, 0.02, size=(3000, 3))).astype(np.float32)
    ref = brute64(p)
    got = knn.distCUDA2(torch.from_numpy(p).cuda()).cpu().numpy()
    assert (ref > 0).all()
    np.testing.assert_allclose(got, ref, rtol=1e-5)
