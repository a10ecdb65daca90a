"""Densification step on top of gs_densify_select / gs_densify_gather: the effect of GaussianModel.densify_and_prune
(/root/reference/scene/gaussian_model.py:1005-1044 -> densify_and_clone :973-1003, densify_and_split :922-971,
densification_postfix :884-920, cat_tensors_to_optimizer :837-881, prune_points :816-835, _prune_optimizer :789-814) on
an optimizer with the reference's six single-tensor groups ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"):
three launches and one host read-back instead of ~150 torch kernels and a dozen read-backs.

The optimizer is edited the way the reference edits it: every group gets a NEW nn.Parameter, its state entry moves to
the new parameter with exp_avg / exp_avg_sq replaced (survivors keep their moments, new Gaussians start at zero) and
"step" untouched.  No CPU path.  NOT yet run on a device (written after round 1's GPU budget was spent); the algorithm is
checked on CPU against the reference's own run (tests/test_densify_oracle.py).
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib

NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
KIND = {"xyz": 1, "scaling": 2}   # 0 copy, 1 position, 2 log-scale, 3 Adam moment (include/grendel_gs_b200.h)


def _as_rows(t):
    """(P, ...) tensor of 4- or 8-byte elements -> (tensor, number of 4-byte elements per row)."""
    if not t.is_cuda or not t.is_contiguous():
        raise TypeError("densification needs contiguous CUDA tensors (no CPU path)")
    if t.element_size() not in (4, 8):
        raise TypeError(f"unsupported element size {t.element_size()}")
    per_row = (t.numel() // max(t.shape[0], 1)) * (t.element_size() // 4)
    return t, per_row


def densify_and_prune(optimizer, xyz_gradient_accum, denom, max_grad, min_opacity, extent, percent_dense, max_screen_size,
                      send_to_gpui_cnt=None, noise=None):
    """-> dict with the six new parameters under the group names, the reset statistics ("xyz_gradient_accum", "denom",
    "max_radii2D", "sum_visible_count_in_one_batch"), "send_to_gpui_cnt" (if given) and "counts" = (kept, clones,
    children per copy, split-selected, new total).  `noise`: optional (>= 2 S, 3) standard-normal draws for the split
    (default: torch.randn on the device, like the reference's torch.normal)."""
    groups = {g["name"]: g for g in optimizer.param_groups}
    if set(groups) != set(NAMES) or any(len(g["params"]) != 1 for g in groups.values()):
        raise ValueError("the optimizer must have the reference's six single-tensor groups " + str(NAMES))
    params = {k: groups[k]["params"][0] for k in NAMES}
    P = params["xyz"].shape[0]
    dev = params["xyz"].device
    if P == 0:
        raise ValueError("no Gaussians")
    stream = torch.cuda.current_stream().cuda_stream
    accum = xyz_gradient_accum.reshape(-1).to(torch.float32).contiguous()
    den = denom.reshape(-1).to(torch.float32).contiguous()
    if accum.numel() != P or den.numel() != P:
        raise ValueError("xyz_gradient_accum / denom must have one entry per Gaussian")
    tb = _lib.query("gs_densify_temp_bytes", P)
    temp = torch.empty((tb,), dtype=torch.uint8, device=dev)
    counts = (C.c_int32 * 6)()
    _lib.call("gs_densify_select", P, accum.data_ptr(), den.data_ptr(), params["scaling"].data_ptr(),
              params["opacity"].data_ptr(), C.c_float(max_grad), C.c_float(min_opacity), C.c_float(extent),
              C.c_float(percent_dense), 1 if max_screen_size else 0, temp.data_ptr(), tb, counts, stream)
    kept, clones, child1, child2, S, new_P = (int(c) for c in counts)
    if noise is None:
        noise = torch.randn((max(2 * S, 1), 3), dtype=torch.float32, device=dev)
    noise = noise.to(device=dev, dtype=torch.float32).contiguous()
    if noise.shape[0] < 2 * S or noise.shape[1] != 3:
        raise ValueError(f"noise must be (>= {2 * S}, 3)")
    # tensor table: parameters, their moments (if the optimizer has state), per-Gaussian bookkeeping
    src, dst, width, kind, outs = [], [], [], [], {}

    def add(name, t, k):
        t, w = _as_rows(t)
        o = torch.empty((new_P,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        src.append(t); dst.append(o); width.append(w); kind.append(k)
        outs[name] = o

    with torch.no_grad():
        for k in NAMES:
            add(k, params[k].detach(), KIND.get(k, 0))
            st = optimizer.state.get(params[k], None)
            if st is not None and "exp_avg" in st:
                add(k + ".exp_avg", st["exp_avg"], 3)
                add(k + ".exp_avg_sq", st["exp_avg_sq"], 3)
        if send_to_gpui_cnt is not None:
            add("send_to_gpui_cnt", send_to_gpui_cnt, 0)
        n = len(src)
        vp, i32 = C.c_void_p * n, C.c_int32 * n
        _lib.call("gs_densify_gather", P, S, new_P, n, vp(*[t.data_ptr() for t in src]), vp(*[t.data_ptr() for t in dst]),
                  i32(*width), i32(*kind), params["scaling"].data_ptr(), params["rotation"].data_ptr(), noise.data_ptr(),
                  temp.data_ptr(), stream)
    # move the optimizer over to the new tensors (gaussian_model.py:789-814 / :837-881)
    result = {}
    for k in NAMES:
        g, old = groups[k], params[k]
        new = nn.Parameter(outs[k].requires_grad_(True))
        st = optimizer.state.pop(old, None)
        if st is not None:
            if "exp_avg" in st:
                st["exp_avg"], st["exp_avg_sq"] = outs[k + ".exp_avg"], outs[k + ".exp_avg_sq"]
            optimizer.state[new] = st
        g["params"][0] = new
        result[k] = new
    result["xyz_gradient_accum"] = torch.zeros((new_P, 1), device=dev)          # densification_postfix :909-914
    result["denom"] = torch.zeros((new_P, 1), device=dev)
    result["max_radii2D"] = torch.zeros((new_P,), device=dev)
    result["sum_visible_count_in_one_batch"] = torch.zeros((new_P,), device=dev)
    if send_to_gpui_cnt is not None:
        result["send_to_gpui_cnt"] = outs["send_to_gpui_cnt"]
    result["counts"] = (kept, clones, child1, S, new_P)
    return result


def append_gaussians(optimizer, new_tensors):
    """cat_tensors_to_optimizer (/root/reference/scene/gaussian_model.py:837-881): rows appended to every parameter, their
    Adam moments start at zero, "step" is kept.  new_tensors: {group name: (n, ...) tensor}.  -> {name: new Parameter}."""
    groups = {g["name"]: g for g in optimizer.param_groups}
    out = {}
    for k in NAMES:
        g = groups[k]
        old, ext = g["params"][0], new_tensors[k].to(device=g["params"][0].device, dtype=g["params"][0].dtype)
        new = nn.Parameter(torch.cat((old.detach(), ext), dim=0).contiguous().requires_grad_(True))
        st = optimizer.state.pop(old, None)
        if st is not None:
            if "exp_avg" in st:
                st["exp_avg"] = torch.cat((st["exp_avg"], torch.zeros_like(ext)), dim=0)
                st["exp_avg_sq"] = torch.cat((st["exp_avg_sq"], torch.zeros_like(ext)), dim=0)
            optimizer.state[new] = st
        g["params"][0] = new
        out[k] = new
    return out
