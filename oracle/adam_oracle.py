"""TEST INFRASTRUCTURE ONLY (see oracle/README or DESIGN.md section 4): CPU restatement of the optimizer step the
reference runs after every batch -- `param.grad /= args.bsz` then `torch.optim.Adam.step()`
(/root/reference/train_internal.py:316-329; optimizer built at /root/reference/scene/gaussian_model.py:257-292 with
eps=1e-15 and six single-tensor groups).  The arithmetic is torch/optim/adam.py's single-tensor path (no weight decay,
no amsgrad, maximize off), restated in numpy float32 one operation at a time.

Parity status: PINNED -- tests/test_optim_oracle.py runs torch.optim.Adam itself (the very class the reference
instantiates) on CPU tensors with the same inputs and compares.
"""
import math

import numpy as np

F = np.float32


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    """One update of one tensor; `step` is the counter AFTER the increment (1 on the first call).
    Returns new (p, m, v) as float32 arrays; inputs are not modified."""
    p, g, m, v = (np.asarray(a, F) for a in (p, g, m, v))
    g = g * F(grad_scale)
    # exp_avg.lerp_(grad, 1 - beta1)
    m = m + F(1.0 - beta1) * (g - m)
    # exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    v = v * F(beta2) + F(1.0 - beta2) * (g * g)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    # denom = (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps); param.addcdiv_(exp_avg, denom, value=-step_size)
    denom = np.sqrt(v) / F(math.sqrt(bc2)) + F(eps)
    p = p - F(step_size) * (m / denom)
    return p.astype(F), m.astype(F), v.astype(F)
