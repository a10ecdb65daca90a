"""CPU: the Adam restatement (oracle/adam_oracle.py) against torch.optim.Adam itself -- the class the reference
instantiates (scene/gaussian_model.py:292) -- on the reference's six parameter groups, learning rates and eps."""
import numpy as np
import torch

from oracle import adam_oracle

# name, trailing shape, lr (arguments/__init__.py defaults: position_lr_init 0.00016 (x spatial scale), feature_lr 0.0025,
# feature_lr / 20, opacity_lr 0.05, scaling_lr 0.005, rotation_lr 0.001)
GROUPS = [("xyz", (3,), 0.00016 * 4.2), ("f_dc", (1, 3), 0.0025), ("f_rest", (15, 3), 0.0025 / 20.0),
          ("opacity", (1,), 0.05), ("scaling", (3,), 0.005), ("rotation", (4,), 0.001)]


def make(n, seed, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    params = [torch.randn((n,) + shp, generator=g).to(device).requires_grad_(True) for _, shp, _ in GROUPS]
    groups = [{"params": [p], "lr": lr, "name": name} for p, (name, _, lr) in zip(params, GROUPS)]
    return params, groups


def grads_for(params, step, scale):
    g = torch.Generator().manual_seed(1000 + step)
    return [(torch.randn(p.shape, generator=g) * scale).to(p.device) for p in params]


def run_oracle(n, steps, bsz, betas=(0.9, 0.999), eps=1e-15, skip=None):
    params, _ = make(n, 3)
    P = [p.detach().numpy().copy() for p in params]
    M = [np.zeros_like(a) for a in P]
    V = [np.zeros_like(a) for a in P]
    count = [0] * len(P)
    for s in range(1, steps + 1):
        gs = grads_for(params, s, scale=10.0 ** (s % 3 - 2))
        for k, (_, _, lr) in enumerate(GROUPS):
            if skip is not None and skip(s, k):
                continue
            count[k] += 1
            P[k], M[k], V[k] = adam_oracle.adam_step(P[k], gs[k].numpy(), M[k], V[k], count[k], lr, betas[0], betas[1], eps,
                                                     grad_scale=1.0 / bsz)
    return P, M, V


def check(got, ref, lr, what):
    """m and v within a few fp32 ulps; p within a few ulps of the update size."""
    (p, m, v), (rp, rm, rv) = got, ref
    # exp_avg is a signed running mean: entries that nearly cancel carry the absolute rounding of their larger terms
    np.testing.assert_allclose(m, rm, rtol=2e-6, atol=2e-7 * float(np.abs(rm).max()), err_msg=what + " exp_avg")
    np.testing.assert_allclose(v, rv, rtol=2e-6, atol=1e-20, err_msg=what + " exp_avg_sq")
    # p: the update itself is ~lr and agrees to ~1e-5 of that; the final rounding of p can flip by an ulp per step
    err = np.abs(p - rp)
    assert (err <= 4e-5 * lr + 6e-7 * np.abs(rp)).all(), (what, float(err.max()))


def test_adam_oracle_matches_torch_adam_on_cpu():
    n, steps, bsz = 257, 7, 4
    skip = lambda s, k: (s == 3 and k == 2) or (s == 5 and k == 0)   # a parameter whose .grad is None that step
    P, M, V = run_oracle(n, steps, bsz, skip=skip)
    params, groups = make(n, 3)
    opt = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    for s in range(1, steps + 1):
        gs = grads_for(params, s, scale=10.0 ** (s % 3 - 2))
        for k, p in enumerate(params):
            p.grad = None if skip(s, k) else gs[k].clone()
        for p in params:                       # train_internal.py:319-324
            if p.grad is not None:
                p.grad /= bsz
        opt.step()
        opt.zero_grad(set_to_none=True)
    for k, (name, _, lr) in enumerate(GROUPS):
        st = opt.state[params[k]]
        check((P[k], M[k], V[k]), (params[k].detach().numpy(), st["exp_avg"].numpy(), st["exp_avg_sq"].numpy()), lr, name)
        assert int(st["step"]) == steps - (1 if name in ("f_rest", "xyz") else 0)


def test_fused_adam_has_no_cpu_path_and_torch_state_layout():
    import pytest
    from gs_b200.optim import FusedAdam
    params, groups = make(8, 1)
    opt = FusedAdam(groups, lr=0.0, eps=1e-15)
    assert [g["name"] for g in opt.param_groups] == [g[0] for g in GROUPS]
    assert all(g["eps"] == 1e-15 and tuple(g["betas"]) == (0.9, 0.999) for g in opt.param_groups)
    opt.step()                                   # no gradients yet: nothing to do, no library call
    params[0].grad = torch.ones_like(params[0])
    with pytest.raises(TypeError):               # CPU tensors are rejected, never updated on the host
        opt.step()
    ref_params, ref_groups = make(8, 1)
    ref = torch.optim.Adam(ref_groups, lr=0.0, eps=1e-15)
    assert opt.state_dict()["param_groups"][0].keys() >= {"lr", "betas", "eps", "params", "name"}
    # a checkpoint written by the reference's optimizer (with state) loads into FusedAdam ...
    for p in ref_params:
        p.grad = torch.ones_like(p)
    ref.step()
    opt.load_state_dict(ref.state_dict())
    st = opt.state[opt.param_groups[2]["params"][0]]
    assert int(st["step"]) == 1 and st["exp_avg"].shape == ref_params[2].shape
    # ... and one written by FusedAdam loads into torch.optim.Adam, which can then step (load_state_dict replaces the
    # groups by the saved ones, so they must carry every key torch's step reads)
    back = torch.optim.Adam(make(8, 1)[1], lr=0.0, eps=1e-15)
    back.load_state_dict(opt.state_dict())
    for g in back.param_groups:
        for p in g["params"]:
            p.grad = torch.ones_like(p)
    back.step()
    assert int(back.state[back.param_groups[0]["params"][0]]["step"]) == 2


def test_adam_oracle_sqrt_lr_scaling_betas():
    """lr_scale_mode 'sqrt' rewrites eps and betas per group (scene/gaussian_model.py:301-309)."""
    bsz = 4
    betas, eps = (0.9 ** bsz, 0.999 ** bsz), 1e-15 / np.sqrt(bsz)
    P, M, V = run_oracle(64, 4, 1, betas=betas, eps=eps)
    params, groups = make(64, 3)
    opt = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    for gr in opt.param_groups:
        gr["eps"] /= np.sqrt(bsz)
        gr["betas"] = [b ** bsz for b in gr["betas"]]
    for s in range(1, 5):
        for p, g in zip(params, grads_for(params, s, scale=10.0 ** (s % 3 - 2))):
            p.grad = g
        opt.step()
    for k, (name, _, lr) in enumerate(GROUPS):
        st = opt.state[params[k]]
        check((P[k], M[k], V[k]), (params[k].detach().numpy(), st["exp_avg"].numpy(), st["exp_avg_sq"].numpy()), lr, name)
