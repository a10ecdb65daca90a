"""Generates tests/golden/densify.npz by running the REFERENCE's own densification code
(/root/reference/scene/gaussian_model.py:789-1044: densify_and_prune -> densify_and_clone, densify_and_split,
densification_postfix / cat_tensors_to_optimizer, prune_points / _prune_optimizer) on CPU tensors.

Only runs in the authoring container (needs /root/reference).  The reference hard-codes device="cuda" in a few
allocations; they are redirected to the CPU here, and torch.normal is replaced by `mean + std * NOISE` with a recorded
standard-normal NOISE so that the restatement (oracle/densify_oracle.py) and the CUDA path can consume the same draws.
Nothing of the reference is copied: its functions are CALLED, their inputs and outputs are stored.

    python tests/golden/make_densify_golden.py
"""
import io
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "grendel-gs_b200"), os.path.join(ROOT, "grendel-gs_b200", "shims"), REF]

_zeros = torch.zeros


def zeros_cpu(*a, **k):
    k.pop("device", None)
    return _zeros(*a, **k)


torch.zeros = zeros_cpu
torch.cuda.empty_cache = lambda: None
NOISE = {"z": None}
torch.normal = lambda mean, std: mean + std * NOISE["z"][: std.shape[0]].to(std.dtype)

import utils.general_utils as utils  # noqa: E402  (the reference's)
import scene.gaussian_model as gm    # noqa: E402

NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
SHAPES = {"xyz": (3,), "f_dc": (1, 3), "f_rest": (15, 3), "opacity": (1,), "scaling": (3,), "rotation": (4,)}


def build(P, seed, world=2):
    g = torch.Generator().manual_seed(seed)
    m = gm.GaussianModel(3)
    extent, pd = 5.0, 0.01
    t = {"xyz": torch.randn((P, 3), generator=g) * 2.0,
         "f_dc": torch.randn((P, 1, 3), generator=g),
         "f_rest": torch.randn((P, 15, 3), generator=g) * 0.1,
         "opacity": torch.randn((P, 1), generator=g) * 2.5,
         "scaling": torch.randn((P, 3), generator=g) * 1.0 + float(np.log(pd * extent)),
         "rotation": torch.randn((P, 4), generator=g)}
    m._xyz, m._features_dc, m._features_rest = (torch.nn.Parameter(t[k].clone()) for k in ("xyz", "f_dc", "f_rest"))
    m._opacity, m._scaling, m._rotation = (torch.nn.Parameter(t[k].clone()) for k in ("opacity", "scaling", "rotation"))
    params = dict(xyz=m._xyz, f_dc=m._features_dc, f_rest=m._features_rest, opacity=m._opacity, scaling=m._scaling,
                  rotation=m._rotation)
    m.optimizer = torch.optim.Adam([{"params": [params[k]], "lr": 1e-3, "name": k} for k in NAMES], lr=0.0, eps=1e-15)
    for _ in range(2):   # populate exp_avg / exp_avg_sq / step
        for k in NAMES:
            params[k].grad = torch.randn(params[k].shape, generator=g) * 0.01
        m.optimizer.step()
    m.optimizer.zero_grad(set_to_none=True)
    m.percent_dense = pd
    m.xyz_gradient_accum = torch.rand((P, 1), generator=g) * 0.0006 * 3
    m.denom = torch.randint(0, 7, (P, 1), generator=g).float()       # zeros -> NaN gradients -> treated as 0
    m.xyz_gradient_accum = m.xyz_gradient_accum * m.denom / 3.0 + (m.denom == 0) * 0.0
    m.max_radii2D = torch.zeros((P,))
    m.sum_visible_count_in_one_batch = torch.zeros((P,))
    m.send_to_gpui_cnt = torch.randint(0, 50, (P, world), generator=g)
    return m, extent


def snapshot(m):
    out = {}
    for gidx, grp in enumerate(m.optimizer.param_groups):
        p = grp["params"][0]
        st = m.optimizer.state[p]
        out[grp["name"]] = p.detach().numpy().copy()
        out[grp["name"] + ".exp_avg"] = st["exp_avg"].numpy().copy()
        out[grp["name"] + ".exp_avg_sq"] = st["exp_avg_sq"].numpy().copy()
    out["send_to_gpui_cnt"] = m.send_to_gpui_cnt.numpy().copy()
    out["xyz_gradient_accum"] = m.xyz_gradient_accum.numpy().copy()
    out["denom"] = m.denom.numpy().copy()
    return out


def main():
    utils.set_args(SimpleNamespace(gaussians_distribution=True))
    utils.set_log_file(io.StringIO())
    utils.DEFAULT_GROUP = utils.SingleGPUGroup()
    blob = {}
    for case, (P, seed, max_screen) in enumerate([(400, 11, 20), (257, 12, None)]):
        m, extent = build(P, seed)
        g = torch.Generator().manual_seed(100 + seed)
        NOISE["z"] = torch.randn((2 * P, 3), generator=g)
        before = snapshot(m)
        max_grad, min_opacity = 0.0002, 0.005
        m.densify_and_prune(max_grad, min_opacity, extent, max_screen)
        after = snapshot(m)
        for k, v in before.items():
            blob[f"c{case}.in.{k}"] = v
        for k, v in after.items():
            blob[f"c{case}.out.{k}"] = v
        blob[f"c{case}.noise"] = NOISE["z"].numpy()
        blob[f"c{case}.scalars"] = np.array([max_grad, min_opacity, extent, m.percent_dense, 1.0 if max_screen else 0.0])
        print(f"case {case}: {P} -> {after['xyz'].shape[0]} Gaussians")
    np.savez_compressed(os.path.join(HERE, "densify.npz"), **blob)


if __name__ == "__main__":
    main()
