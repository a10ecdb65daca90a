"""-m gpu:
gs_densify_select / gs_densify_gather through gs_b200.densify.densify_and_prune against tests/golden/densify.npz, i.e.
against the REFERENCE's own densify_and_prune run (tests/golden/make_densify_golden.py), with the same normal draws."""
import os

import numpy as np
import pytest
import torch

from test_densify_oracle import load

pytestmark = pytest.mark.gpu
NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


@pytest.mark.parametrize("case", [0, 1])
def test_fused_densification_matches_the_reference_run(case):
    from gs_b200 import densify
    ins, outs, noise, (max_grad, min_opacity, extent, pd, screen) = load(case)
    dev = "cuda"
    params = {k: torch.nn.Parameter(torch.from_numpy(ins[k]).to(dev)) for k in NAMES}
    opt = torch.optim.Adam([{"params": [params[k]], "lr": 1e-3, "name": k} for k in NAMES], lr=0.0, eps=1e-15)
    for k in NAMES:
        opt.state[params[k]] = {"step": torch.tensor(2.0), "exp_avg": torch.from_numpy(ins[k + ".exp_avg"]).to(dev),
                                "exp_avg_sq": torch.from_numpy(ins[k + ".exp_avg_sq"]).to(dev)}
    res = densify.densify_and_prune(opt, torch.from_numpy(ins["xyz_gradient_accum"]).to(dev),
                                    torch.from_numpy(ins["denom"]).to(dev), float(max_grad), float(min_opacity), float(extent),
                                    float(pd), 20 if screen else None,
                                    send_to_gpui_cnt=torch.from_numpy(ins["send_to_gpui_cnt"]).to(dev),
                                    noise=torch.from_numpy(noise).to(dev))
    new_P = outs["xyz"].shape[0]
    assert res["counts"][-1] == new_P
    for k in NAMES:
        p = opt.param_groups[NAMES.index(k)]["params"][0]
        assert p is res[k] and p.requires_grad and p.shape[0] == new_P
        got = {k: p.detach().cpu().numpy(), k + ".exp_avg": opt.state[p]["exp_avg"].cpu().numpy(),
               k + ".exp_avg_sq": opt.state[p]["exp_avg_sq"].cpu().numpy()}
        assert float(opt.state[p]["step"]) == 2.0
        for name, v in got.items():
            if name in ("xyz", "scaling"):
                np.testing.assert_allclose(v, outs[name], rtol=2e-6, atol=2e-6, err_msg=name)
            else:
                assert np.array_equal(v, outs[name]), name
    assert np.array_equal(res["send_to_gpui_cnt"].cpu().numpy(), outs["send_to_gpui_cnt"])
    assert res["xyz_gradient_accum"].shape == (new_P, 1) and float(res["denom"].abs().sum()) == 0.0
    # the optimizer keeps working on the new tensors
    for k in NAMES:
        res[k].grad = torch.ones_like(res[k])
    opt.step()
