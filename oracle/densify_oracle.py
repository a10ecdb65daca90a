"""TEST INFRASTRUCTURE ONLY: numpy restatement of the reference's densification step
(/root/reference/scene/gaussian_model.py:1005-1044 densify_and_prune, :973-1003 densify_and_clone, :922-971
densify_and_split, :884-920 densification_postfix, :837-881 cat_tensors_to_optimizer, :816-835 prune_points, :789-814
_prune_optimizer; rotation matrix of /root/reference/utils/general_utils.py:416-438).

Parity status: PINNED -- tests/golden/densify.npz holds inputs and outputs of the reference's own functions run on CPU
tensors (tests/golden/make_densify_golden.py); tests/test_densify_oracle.py compares this restatement with them.

State is a dict of float32 arrays: the six raw parameters under the reference's group names (xyz, f_dc, f_rest, opacity,
scaling, rotation), their Adam moments under "<name>.exp_avg" / "<name>.exp_avg_sq", and the per-Gaussian bookkeeping
arrays send_to_gpui_cnt (int), xyz_gradient_accum, denom.  `noise` is the (>= 2 S, 3) block of standard-normal draws the
split consumes (the reference draws them with torch.normal; the generator stream is not reproducible across devices, so
every implementation takes the draws as an input).
"""
import numpy as np

F = np.float32
NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def build_rotation(r):
    """(n,4) raw quaternions (w,x,y,z) -> (n,3,3), utils/general_utils.py:416-438."""
    r = r.astype(F)
    norm = np.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.zeros((r.shape[0], 3, 3), F)
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - w * z)
    R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y)
    R[:, 2, 1] = 2 * (y * z + w * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _append(st, new, new_send):
    """densification_postfix + cat_tensors_to_optimizer: parameters grow by `new`, moments by zeros, statistics reset."""
    for k in NAMES:
        st[k] = np.concatenate([st[k], new[k]], axis=0)
        for m in (".exp_avg", ".exp_avg_sq"):
            st[k + m] = np.concatenate([st[k + m], np.zeros_like(new[k])], axis=0)
    n = st["xyz"].shape[0]
    st["xyz_gradient_accum"] = np.zeros((n, 1), F)
    st["denom"] = np.zeros((n, 1), F)
    st["send_to_gpui_cnt"] = np.concatenate([st["send_to_gpui_cnt"], new_send], axis=0)


def _prune(st, mask):
    """prune_points: drop the rows where mask is True from every per-Gaussian array."""
    keep = ~mask
    for k in list(st):
        st[k] = st[k][keep]


def densify_and_prune(state, noise, max_grad, min_opacity, extent, percent_dense, use_screen_size, N=2):
    st = {k: np.array(v, copy=True) for k, v in state.items()}
    with np.errstate(divide="ignore", invalid="ignore"):
        grads = (st["xyz_gradient_accum"].astype(F) / st["denom"].astype(F)).astype(F)
    grads[np.isnan(grads)] = 0.0
    # ---- densify_and_clone --------------------------------------------------------------------------------------
    scale = np.exp(st["scaling"].astype(F))
    sel = (np.linalg.norm(grads, axis=-1) >= F(max_grad)) & (scale.max(axis=1) <= F(percent_dense * extent))
    _append(st, {k: st[k][sel] for k in NAMES}, st["send_to_gpui_cnt"][sel])
    n_clone = int(sel.sum())
    # ---- densify_and_split --------------------------------------------------------------------------------------
    n_init = st["xyz"].shape[0]
    padded = np.zeros((n_init,), F)
    padded[: grads.shape[0]] = grads.reshape(-1)
    scale = np.exp(st["scaling"].astype(F))
    sel = (padded >= F(max_grad)) & (scale.max(axis=1) > F(percent_dense * extent))
    S = int(sel.sum())
    stds = np.tile(scale[sel], (N, 1))
    samples = (stds * noise[: N * S].astype(F)).astype(F)
    rots = np.tile(build_rotation(st["rotation"][sel]), (N, 1, 1))
    new = {"xyz": (np.einsum("nij,nj->ni", rots, samples).astype(F) + np.tile(st["xyz"][sel], (N, 1))).astype(F),
           "scaling": np.log(np.tile(scale[sel], (N, 1)) / F(0.8 * N)).astype(F),
           "rotation": np.tile(st["rotation"][sel], (N, 1)),
           "f_dc": np.tile(st["f_dc"][sel], (N, 1, 1)), "f_rest": np.tile(st["f_rest"][sel], (N, 1, 1)),
           "opacity": np.tile(st["opacity"][sel], (N, 1))}
    _append(st, new, np.tile(st["send_to_gpui_cnt"][sel], (N, 1)))
    _prune(st, np.concatenate([sel, np.zeros((N * S,), bool)]))
    # ---- final prune ------------------------------------------------------------------------------------------------
    opacity = (F(1) / (F(1) + np.exp(-st["opacity"].astype(F)))).astype(F)
    mask = (opacity < F(min_opacity)).reshape(-1)
    if use_screen_size:   # max_radii2D is all zero in the reference (its own comment, :1027-1036): only the world-size test acts
        mask = mask | (np.exp(st["scaling"].astype(F)).max(axis=1) > F(0.1 * extent))
    _prune(st, mask)
    st["_counts"] = np.array([n_clone, S, int(mask.sum())])
    return st
