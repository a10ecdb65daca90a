"""CPU: the CPU-tensor path of simple_knn._C.distCUDA2 against a float64 brute force, incl. the off-origin dense cloud on
which the matmul form of torch.cdist loses every digit."""
import numpy as np
import torch


def brute64(p):
    q = p.astype(np.float64)
    d = ((q[:, None, :] - q[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d, np.inf)
    k = min(3, len(p) - 1)
    return np.sort(d, axis=1)[:, :k].mean(axis=1) if k > 0 else np.zeros(len(p))


def test_cpu_path_is_exact_on_an_off_origin_dense_cloud():
    from simple_knn._C import distCUDA2
    rng = np.random.default_rng(7)
    p = (np.array([30.0, -20.0, 15.0]) + rng.uniform(-0.02, 0.02, size=(3000, 3))).astype(np.float32)
    ref = brute64(p)
    assert (ref > 0).all()
    np.testing.assert_allclose(distCUDA2(torch.from_numpy(p)).numpy(), ref, rtol=1e-5)
    bad = torch.cdist(torch.from_numpy(p), torch.from_numpy(p)).pow(2).topk(4, dim=1, largest=False).values[:, 1:].mean(1)
    assert float((bad.numpy() - ref).__abs__().max() / ref.max()) > 0.1     # what the old shim returned


def test_cpu_path_small_and_duplicates():
    from simple_knn._C import distCUDA2
    for n in (1, 2, 3, 4, 50):
        p = np.random.default_rng(n).normal(size=(n, 3)).astype(np.float32)
        if n > 10:
            p[5] = p[7]
        np.testing.assert_allclose(distCUDA2(torch.from_numpy(p)).numpy(), brute64(p), rtol=1e-5, atol=1e-12)
