"""Import shim for `from simple_knn._C import distCUDA2` (/root/reference/scene/gaussian_model.py:20,163-166).
Init-time only (scale initialisation); implemented with torch ops in chunks."""
import torch


def distCUDA2(points):
    """Mean squared distance to the 3 nearest neighbours of every point, (N,) float32."""
    pts = points.float()
    n = pts.shape[0]
    out = torch.empty((n,), dtype=torch.float32, device=pts.device)
    chunk = max(1, min(n, (1 << 26) // max(n, 1)))
    for s in range(0, n, chunk):
        d = torch.cdist(pts[s:s + chunk], pts).pow(2)
        k = min(4, n)
        vals = d.topk(k, dim=1, largest=False).values[:, 1:]
        out[s:s + chunk] = vals.mean(dim=1) if k > 1 else 0.0
    return out
