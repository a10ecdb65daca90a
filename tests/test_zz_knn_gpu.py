"""-m gpu: gs_knn3_mean_dist2 against a numpy brute
force (the definition: exact 3 nearest OTHER points, duplicates count at distance 0) and against the torch shim."""
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
    if n > 3:
        np.testing.assert_allclose(got, knn._dist2_torch(t).cpu().numpy(), rtol=1e-3, atol=1e-5)   # cdist goes through sqrt
