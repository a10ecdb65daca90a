"""-m gpu: gs_adam_step (through gs_b200.optim.FusedAdam) against the CPU restatement (oracle/adam_oracle.py, itself
pinned to torch.optim.Adam on CPU) and against torch.optim.Adam running on the same device."""
import copy

import numpy as np
import pytest
import torch

from test_optim_oracle import GROUPS, check, grads_for, make, run_oracle

pytestmark = pytest.mark.gpu


def _state(opt, p):
    st = opt.state[p]
    return p.detach().cpu().numpy(), st["exp_avg"].cpu().numpy(), st["exp_avg_sq"].cpu().numpy()


def test_fused_adam_matches_oracle_and_torch_adam():
    from gs_b200.optim import FusedAdam
    n, steps, bsz = 20011, 6, 4     # 60033 / 900495 / 20011-element tensors: float4 bodies with scalar tails
    skip = lambda s, k: (s == 2 and k == 1) or (s == 4 and k == 5)
    P, M, V = run_oracle(n, steps, bsz, skip=skip)
    params_f, groups_f = make(n, 3, "cuda")
    params_t, groups_t = make(n, 3, "cuda")
    fused = FusedAdam(groups_f, lr=0.0, eps=1e-15)
    ref = torch.optim.Adam(groups_t, lr=0.0, eps=1e-15)
    for s in range(1, steps + 1):
        gs = grads_for(params_f, s, scale=10.0 ** (s % 3 - 2))
        for k in range(len(GROUPS)):
            params_f[k].grad = None if skip(s, k) else gs[k].clone()
            params_t[k].grad = None if skip(s, k) else gs[k].clone() / bsz     # train_internal.py:319-324
        fused.step(grad_scale=1.0 / bsz)      # the division by bsz happens inside the kernel
        ref.step()
        fused.zero_grad(set_to_none=True)
        ref.zero_grad(set_to_none=True)
    for k, (name, _, lr) in enumerate(GROUPS):
        got = _state(fused, params_f[k])
        check(got, (P[k], M[k], V[k]), lr, "oracle " + name)
        check(got, _state(ref, params_t[k]), lr, "torch.optim.Adam " + name)
        assert int(fused.state[params_f[k]]["step"]) == int(ref.state[params_t[k]]["step"])


def test_fused_adam_state_is_interchangeable_with_torch_adam():
    """Checkpoints and the reference's direct edits of optimizer.state (gaussian_model.py:771-881) see torch's layout:
    a FusedAdam state_dict loads into torch.optim.Adam (and back) and both continue identically."""
    from gs_b200.optim import FusedAdam
    n = 4099
    params_f, groups_f = make(n, 5, "cuda")
    params_t, groups_t = make(n, 5, "cuda")
    fused = FusedAdam(groups_f, lr=0.0, eps=1e-15)
    for s in (1, 2):
        for p, g in zip(params_f, grads_for(params_f, s, 0.1)):
            p.grad = g
        fused.step()
    with torch.no_grad():
        for a, b in zip(params_t, params_f):
            a.copy_(b)
    ref = torch.optim.Adam(groups_t, lr=0.0, eps=1e-15)
    # load_state_dict keeps tensors whose dtype/device already match (no copy): deepcopy, or the optimizers would
    # share exp_avg / exp_avg_sq / step and both step() calls would advance the same moments
    ref.load_state_dict(copy.deepcopy(fused.state_dict()))
    back = FusedAdam(groups_f, lr=0.0, eps=1e-15)
    back.load_state_dict(copy.deepcopy(ref.state_dict()))
    for p, q, g in zip(params_f, params_t, grads_for(params_f, 3, 0.1)):
        p.grad, q.grad = g.clone(), g.clone()
    back.step()
    ref.step()
    for k, (name, _, lr) in enumerate(GROUPS):
        check(_state(back, params_f[k]), _state(ref, params_t[k]), lr, name)
        assert int(back.state[params_f[k]]["step"]) == 3


def test_fused_adam_unaligned_and_rejected_inputs():
    from gs_b200.optim import FusedAdam
    base = torch.randn((1001,), device="cuda")
    p = base[1:].detach().requires_grad_(True)            # 4-byte aligned only: the scalar path of the kernel
    q = p.detach().clone().requires_grad_(True)
    g = torch.randn((1000,), device="cuda")
    p.grad, q.grad = g.clone(), g.clone()
    a, b = FusedAdam([p], lr=0.01), torch.optim.Adam([q], lr=0.01)
    a.step()
    b.step()
    check(_state(a, p), _state(b, q), 0.01, "unaligned")
    cpu = torch.zeros((4,), requires_grad=True)
    cpu.grad = torch.ones((4,))
    with pytest.raises(TypeError):
        FusedAdam([cpu], lr=0.01).step()
