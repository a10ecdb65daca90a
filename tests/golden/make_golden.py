"""Generates tests/golden/*.npz by IMPORTING the reference's own Python (run in the authoring
container only; /root/reference does not exist on the GPU box).  These are the only pieces of
the hot path the reference keeps as runnable CPU Python:
  utils/sh_utils.py:eval_sh            -> SH basis / sign convention
  utils/graphics_utils.py              -> getWorld2View2, getProjectionMatrix
  scene/cameras.py:84-100 (restated)   -> transposed storage, full_proj, camera_center
  utils/loss_utils.py                  -> pixelwise_l1_with_mask, pixelwise_ssim_with_mask
Usage: python tests/golden/make_golden.py
"""
import importlib.util
import math
import os

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    sh_utils = load("utils/sh_utils.py", "ref_sh_utils")
    gfx = load("utils/graphics_utils.py", "ref_graphics_utils")
    loss_utils = load("utils/loss_utils.py", "ref_loss_utils")
    rng = np.random.default_rng(1234)

    # --- SH ---------------------------------------------------------------------------
    n = 257
    sh = rng.normal(0, 1, (n, 16, 3)).astype(np.float32)          # operator layout (n,16,3)
    dirs = rng.normal(0, 1, (n, 3))
    dirs = (dirs / np.linalg.norm(dirs, axis=1, keepdims=True)).astype(np.float32)
    out = {}
    for deg in range(4):
        # eval_sh wants [..., C, coeffs]
        res = sh_utils.eval_sh(deg, torch.tensor(sh).transpose(1, 2), torch.tensor(dirs))
        out[f"rgb_deg{deg}"] = res.numpy()
    np.savez(os.path.join(HERE, "sh.npz"), sh=sh, dirs=dirs, **out)

    # --- cameras ----------------------------------------------------------------------
    cams = []
    for k in range(6):
        a = rng.uniform(-0.6, 0.6, 3)
        Rx = np.array([[1, 0, 0], [0, math.cos(a[0]), -math.sin(a[0])], [0, math.sin(a[0]), math.cos(a[0])]])
        Ry = np.array([[math.cos(a[1]), 0, math.sin(a[1])], [0, 1, 0], [-math.sin(a[1]), 0, math.cos(a[1])]])
        Rz = np.array([[math.cos(a[2]), -math.sin(a[2]), 0], [math.sin(a[2]), math.cos(a[2]), 0], [0, 0, 1]])
        R = Rx @ Ry @ Rz
        T = rng.uniform(-2, 2, 3)
        trans = rng.uniform(-1, 1, 3) if k % 2 else np.zeros(3)
        scale = 1.0 if k < 3 else 1.7
        fovx, fovy = rng.uniform(0.6, 1.4), rng.uniform(0.5, 1.2)
        wv = torch.tensor(gfx.getWorld2View2(R, T, trans, scale)).transpose(0, 1)
        pm = gfx.getProjectionMatrix(znear=0.01, zfar=100.0, fovX=fovx, fovY=fovy).transpose(0, 1)
        full = (wv.unsqueeze(0).bmm(pm.unsqueeze(0))).squeeze(0)
        center = wv.inverse()[3, :3]
        cams.append(dict(R=R, T=T, trans=trans, scale=scale, fovx=fovx, fovy=fovy, world_view=wv.numpy(),
                         proj=pm.numpy(), full=full.numpy(), center=center.numpy()))
    np.savez(os.path.join(HERE, "cameras.npz"), **{f"{k}_{i}": np.asarray(c[k]) for i, c in enumerate(cams) for k in c})

    # --- loss -------------------------------------------------------------------------
    img = torch.tensor(rng.uniform(0, 1, (3, 45, 70)).astype(np.float32))
    gt_u8 = torch.tensor(rng.integers(0, 256, (3, 45, 70), dtype=np.uint8))
    gt = torch.clamp(gt_u8 / 255.0, 0.0, 1.0)
    mask = torch.ones((45, 70), dtype=torch.bool)
    x = img.clone().requires_grad_(True)
    n_pix = 200 * 70  # the strip belongs to a taller full image
    l1 = loss_utils.pixelwise_l1_with_mask(x, gt, mask).sum() / (n_pix * 3)
    ss = loss_utils.pixelwise_ssim_with_mask(x, gt, mask).sum() / (n_pix * 3)
    loss = (1.0 - 0.2) * l1 + 0.2 * (1.0 - ss)
    loss.backward()
    np.savez(os.path.join(HERE, "loss.npz"), img=img.numpy(), gt_u8=gt_u8.numpy(), n_pix=n_pix, l1=l1.item(),
             ssim=ss.item(), loss=loss.item(), grad=x.grad.numpy())
    # --- covariance: build_rotation / build_scaling_rotation / strip_symmetric (utils/general_utils.py:400-451) chained as
    # build_covariance_from_scaling_rotation does (scene/gaussian_model.py:35-39).  The functions hard-code device="cuda";
    # they are executed here from the reference's own source text with that one literal replaced by "cpu".
    src = open(os.path.join(REF, "utils/general_utils.py")).read()
    a, b = src.index("def strip_lowerdiag"), src.index("def safe_state")
    ns = {"torch": torch}
    exec(src[a:b].replace('device="cuda"', 'device="cpu"'), ns)
    n = 300
    scal = np.exp(rng.normal(-2.0, 1.0, (n, 3))).astype(np.float32)
    rot = rng.normal(0, 1, (n, 4)).astype(np.float32)              # unnormalised, as the raw parameter
    mod = 1.0
    L = ns["build_scaling_rotation"](mod * torch.tensor(scal), torch.tensor(rot))
    cov = ns["strip_symmetric"](L @ L.transpose(1, 2))
    np.savez(os.path.join(HERE, "cov3d.npz"), scales=scal, rotations=rot, R=ns["build_rotation"](torch.tensor(rot)).numpy(),
             cov6=cov.numpy())
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
