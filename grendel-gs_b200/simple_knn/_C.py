"""`from simple_knn._C import distCUDA2` (/root/reference/scene/gaussian_model.py:20,163-166): start-up scale
initialisation only, not on the training path.

CUDA tensors go to the C-ABI kernel gs_knn3_mean_dist2 (exact tiled brute force, self excluded by index; validated on a
B200 by tests/test_zz_knn_gpu.py) and fail loudly if the library is missing.  CPU tensors (the reference never passes
one: gaussian_model.py:163 calls `.cuda()` first) get the same definition from exact coordinate differences -- NOT from
torch.cdist's |a|^2 + |b|^2 - 2ab matmul form, which cancels catastrophically in fp32 for near neighbours of an
off-origin cloud (every distance of a 3000-point cloud with 2.6e-3 spacing around (30,-20,15) came out 0)."""
import torch


def _dist2_exact_torch(pts):
    """Exact 3-NN mean squared distance from coordinate differences, in chunks; self excluded BY INDEX."""
    n = pts.shape[0]
    out = torch.zeros((n,), dtype=torch.float32, device=pts.device)
    k = min(3, n - 1)
    if k <= 0:
        return out
    chunk = max(1, min(n, (1 << 24) // max(n, 1)))
    cols = torch.arange(n, device=pts.device)
    for s in range(0, n, chunk):
        q = pts[s:s + chunk]
        d = (q[:, None, :] - pts[None, :, :]).square().sum(-1)
        d[torch.arange(q.shape[0], device=pts.device), cols[s:s + chunk]] = float("inf")
        out[s:s + chunk] = d.topk(k, dim=1, largest=False).values.mean(dim=1)
    return out


def _dist2_kernel(pts):
    from gs_b200 import _lib
    if not pts.is_cuda:
        raise TypeError("distCUDA2 kernel path needs a CUDA tensor")
    pts = pts.contiguous()
    out = torch.empty((pts.shape[0],), dtype=torch.float32, device=pts.device)
    _lib.call("gs_knn3_mean_dist2", pts.shape[0], pts.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return out


def distCUDA2(points):
    """Mean squared distance to the 3 nearest neighbours of every point, (N,) float32."""
    pts = points.float()
    if pts.is_cuda:
        return _dist2_kernel(pts)
    return _dist2_exact_torch(pts)
