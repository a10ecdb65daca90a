"""Differentiable fp64 torch restatement of the operator's FORWARD only (test infrastructure).

Its purpose is to pin the oracle's hand-written backward: torch.autograd differentiates this
forward, and tests/test_oracle.py compares the result with oracle/gs_oracle.c's backward.
The discrete decisions (cull, tile lists, alpha tests, early termination) are taken from the
oracle's forward outputs so both walk exactly the same entries; the straight-through choices
of the published 3DGS backward are reproduced with .detach():
  * alpha = min(0.99, o*G) passes gradient as if unclamped,
  * a view-space x/z (y/z) ratio clamped to the 1.3x guard band is a constant.
"""
import numpy as np
import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def _sh_color(deg, sh, d):
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    res = SH_C0 * sh[:, 0]
    if deg > 0:
        res = res - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = (res + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2 * zz - xx - yy) * sh[:, 6]
               + SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8])
    if deg > 2:
        res = (res + SH_C3[0] * y * (3 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10]
               + SH_C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
               + SH_C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14]
               + SH_C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return res


def preprocess(means3D, scales, rot, shs, opac, cam, scale_modifier=1.0):
    """-> means2D (pixels), conic_opacity, rgb for ALL Gaussians (caller masks by the oracle's radii)."""
    dt = means3D.dtype
    V = torch.as_tensor(cam["viewmatrix"], dtype=dt)
    PM = torch.as_tensor(cam["projmatrix"], dtype=dt)
    campos = torch.as_tensor(cam["campos"], dtype=dt)
    W, H = cam["image_width"], cam["image_height"]
    n = means3D.shape[0]
    ph = torch.cat([means3D, torch.ones(n, 1, dtype=dt)], 1)
    t = ph @ V
    hom = ph @ PM
    pw = 1.0 / (hom[:, 3] + 1e-7)
    ndc = hom[:, :2] * pw[:, None]
    means2D = torch.stack([((ndc[:, 0] + 1) * W - 1) * 0.5, ((ndc[:, 1] + 1) * H - 1) * 0.5], 1)
    r, x, y, z = rot[:, 0], rot[:, 1], rot[:, 2], rot[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(n, 3, 3)
    L = R * (scale_modifier * scales)[:, None, :]
    Sigma = L @ L.transpose(1, 2)
    fx, fy = W / (2 * cam["tanfovx"]), H / (2 * cam["tanfovy"])
    limx, limy = 1.3 * cam["tanfovx"], 1.3 * cam["tanfovy"]
    tz = t[:, 2]
    txtz, tytz = t[:, 0] / tz, t[:, 1] / tz
    cx = torch.where((txtz < -limx) | (txtz > limx), (txtz.clamp(-limx, limx) * tz).detach(), t[:, 0])
    cy = torch.where((tytz < -limy) | (tytz > limy), (tytz.clamp(-limy, limy) * tz).detach(), t[:, 1])
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -fx * cx / (tz * tz), zero, fy / tz, -fy * cy / (tz * tz)], 1).reshape(n, 2, 3)
    Wm = V[:3, :3].T  # world->view rotation, column-vector form
    Tm = J @ Wm
    cov = Tm @ Sigma @ Tm.transpose(1, 2)
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c - b * b
    conic_opacity = torch.stack([c / det, -b / det, a / det, opac[:, 0]], 1)
    d = means3D - campos
    d = d / d.norm(dim=1, keepdim=True)
    rgb = torch.clamp_min(_sh_color(cam["sh_degree"], shs, d) + 0.5, 0.0)
    return means2D, conic_opacity, rgb


def render(means2D, conic_opacity, rgb, bg, H, W, fwd):
    """Composite with the oracle's tile lists / decisions (fwd = Oracle.render_forward output)."""
    dt = means2D.dtype
    gx = (W + 15) // 16
    bg_t = torch.as_tensor(np.asarray(bg), dtype=dt)
    img = torch.zeros(3, H, W, dtype=dt)
    ids_all = torch.as_tensor(fwd["ids"].astype(np.int64))
    ncon = fwd["n_contrib"]
    for t in range(fwd["ranges"].shape[0]):
        if not fwd["compute_locally"][t]:
            continue
        beg, end = int(fwd["ranges"][t, 0]), int(fwd["ranges"][t, 1])
        ty, tx = divmod(t, gx)
        y0, y1, x0, x1 = ty * 16, min(H, ty * 16 + 16), tx * 16, min(W, tx * 16 + 16)
        ys, xs = torch.meshgrid(torch.arange(y0, y1, dtype=dt), torch.arange(x0, x1, dtype=dt), indexing="ij")
        npx = ys.numel()
        if end <= beg:
            img[:, y0:y1, x0:x1] = bg_t[:, None, None]
            continue
        ids = ids_all[beg:end]
        m, co, col = means2D[ids], conic_opacity[ids], rgb[ids]
        dx = m[None, :, 0] - xs.reshape(-1, 1)
        dy = m[None, :, 1] - ys.reshape(-1, 1)
        power = -0.5 * (co[None, :, 0] * dx * dx + co[None, :, 2] * dy * dy) - co[None, :, 1] * dx * dy
        a_raw = co[None, :, 3] * torch.exp(power)
        alpha = a_raw + (a_raw.clamp(max=0.99) - a_raw).detach()
        # entries a pixel blends = those before its last contributor that pass the alpha tests
        last = torch.as_tensor(ncon[y0:y1, x0:x1].astype(np.int64)).reshape(-1, 1)
        k = torch.arange(end - beg).reshape(1, -1)
        live = (k < last) & (power.detach() <= 0) & (alpha.detach() >= 1.0 / 255.0)
        a_eff = torch.where(live, alpha, torch.zeros_like(alpha))
        Texcl = torch.cumprod(torch.cat([torch.ones(npx, 1, dtype=dt), 1 - a_eff[:, :-1]], 1), 1)
        w = a_eff * Texcl
        Tfin = Texcl[:, -1] * (1 - a_eff[:, -1])
        out = w @ col + Tfin[:, None] * bg_t[None, :]
        img[:, y0:y1, x0:x1] = out.T.reshape(3, y1 - y0, x1 - x0)
    return img


def ssim_l1_loss(img, gt, n_pixels_total, lambda_dssim=0.2):
    """The per-strip loss, written with F.conv2d exactly as the live path evaluates it."""
    import torch.nn.functional as F
    g = torch.tensor([np.exp(-((k - 5) ** 2) / (2 * 1.5 ** 2)) for k in range(11)], dtype=torch.float32)
    g = (g / g.sum()).to(img.dtype)
    w2 = (g[:, None] @ g[None, :]).expand(3, 1, 11, 11).contiguous()
    x, y = img[None], gt[None]
    mu1, mu2 = F.conv2d(x, w2, padding=5, groups=3), F.conv2d(y, w2, padding=5, groups=3)
    s1 = F.conv2d(x * x, w2, padding=5, groups=3) - mu1 * mu1
    s2 = F.conv2d(y * y, w2, padding=5, groups=3) - mu2 * mu2
    s12 = F.conv2d(x * y, w2, padding=5, groups=3) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    ssim = m.sum() / (n_pixels_total * 3)
    l1 = (img - gt).abs().sum() / (n_pixels_total * 3)
    return (1 - lambda_dssim) * l1 + lambda_dssim * (1 - ssim), l1, ssim
