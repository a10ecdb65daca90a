"""CPU: oracle/densify_oracle.py against tests/golden/densify.npz -- inputs and outputs of the REFERENCE's own
densify_and_prune (scene/gaussian_model.py:1005-1044 and the functions it calls) run on CPU tensors by
tests/golden/make_densify_golden.py."""
import os

import numpy as np
import pytest

from oracle import densify_oracle as do

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "densify.npz")


def load(case):
    z = np.load(GOLD)
    pre = f"c{case}."
    ins = {k[len(pre) + 3:]: z[k] for k in z.files if k.startswith(pre + "in.")}
    outs = {k[len(pre) + 4:]: z[k] for k in z.files if k.startswith(pre + "out.")}
    return ins, outs, z[pre + "noise"], z[pre + "scalars"]


@pytest.mark.parametrize("case", [0, 1])
def test_densify_oracle_matches_the_reference_run(case):
    ins, outs, noise, (max_grad, min_opacity, extent, pd, screen) = load(case)
    got = do.densify_and_prune(ins, noise, float(max_grad), float(min_opacity), float(extent), float(pd), bool(screen))
    n_clone, n_split, n_pruned = got.pop("_counts")
    assert n_clone > 5 and n_split > 20 and n_pruned > 0           # every branch of the step is exercised
    assert got["xyz"].shape == outs["xyz"].shape
    for k, ref in outs.items():
        assert got[k].shape == ref.shape, k
        if k in ("xyz", "scaling"):   # split children: R(q) (s z) + x and log(s / 1.6): a few ulps between numpy and torch
            np.testing.assert_allclose(got[k], ref, rtol=2e-6, atol=2e-6, err_msg=k)
        else:                         # everything else is copied / zero-filled: exact
            assert np.array_equal(got[k], ref), k
    # moments of the new Gaussians are zero, those of the survivors are carried over
    P0 = ins["xyz"].shape[0]
    assert np.abs(got["f_rest.exp_avg"][: min(P0, 50)]).sum() > 0
    assert np.abs(got["f_rest.exp_avg"][-10:]).sum() == 0
