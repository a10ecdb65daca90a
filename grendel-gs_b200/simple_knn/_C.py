"""`from simple_knn._C import distCUDA2` (/root/reference/scene/gaussian_model.py:20,163-166): start-up scale
initialisation only, not on the training path.

Default: torch ops in chunks (O(N^2) distance blocks; fine for the ~1e5-point COLMAP clouds the reference starts from).
GS_B200_DISTCUDA2=kernel selects the C-ABI kernel gs_knn3_mean_dist2 (exact tiled brute force, ~1 s for 2 M points) --
opt-in until it has been validated on a device (tests/test_zz_knn_gpu.py, GS_B200_EXPERIMENTAL=1)."""
import os

import torch


def _dist2_torch(pts):
    n = pts.shape[0]
    out = torch.empty((n,), dtype=torch.float32, device=pts.device)
    chunk = max(1, min(n, (1 << 26) // max(n, 1)))
    for s in range(0, n, chunk):
        d = torch.cdist(pts[s:s + chunk], pts).pow(2)
        k = min(4, n)
        vals = d.topk(k, dim=1, largest=False).values[:, 1:]
        out[s:s + chunk] = vals.mean(dim=1) if k > 1 else 0.0
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
    if os.environ.get("GS_B200_DISTCUDA2") == "kernel":
        return _dist2_kernel(pts)
    return _dist2_torch(pts)
