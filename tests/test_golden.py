"""Oracle / host helpers against golden vectors generated from the reference's own Python
(tests/golden/make_golden.py).  These pin the SH basis, the camera conventions and the loss."""
import os

import numpy as np

from gs_b200 import synthetic as syn
from oracle.oracle import Oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_sh_basis_matches_reference_eval_sh():
    z = np.load(os.path.join(G, "sh.npz"))
    o = Oracle(np.float32)
    for deg in range(4):
        got = o.eval_sh(deg, z["sh"], z["dirs"])
        np.testing.assert_allclose(got, z[f"rgb_deg{deg}"], rtol=2e-5, atol=2e-6)


def test_camera_matrices_match_reference():
    z = np.load(os.path.join(G, "cameras.npz"))
    for i in range(6):
        w2v = syn.world_to_view(z[f"R_{i}"], z[f"T_{i}"], z[f"trans_{i}"], float(z[f"scale_{i}"]))
        np.testing.assert_allclose(w2v.T, z[f"world_view_{i}"], rtol=0, atol=1e-6)
        pm = syn.projection_matrix(0.01, 100.0, float(z[f"fovx_{i}"]), float(z[f"fovy_{i}"])).T
        assert np.array_equal(pm, z[f"proj_{i}"])
        full = z[f"world_view_{i}"].astype(np.float32) @ pm
        np.testing.assert_allclose(full, z[f"full_{i}"], rtol=1e-6, atol=1e-6)
        center = np.linalg.inv(z[f"world_view_{i}"].astype(np.float64))[3, :3]
        np.testing.assert_allclose(center, z[f"center_{i}"], rtol=1e-5, atol=1e-5)


def test_oracle_projects_like_reference_matrices():
    """A point pushed through the reference's full_proj_transform lands where the oracle says."""
    z = np.load(os.path.join(G, "cameras.npz"))
    o = Oracle(np.float64)
    rng = np.random.default_rng(0)
    for i in range(6):
        V, PM = z[f"world_view_{i}"].astype(np.float64), z[f"full_{i}"].astype(np.float64)
        W, H = 320, 200
        cam = dict(viewmatrix=V, projmatrix=PM, campos=z[f"center_{i}"], image_width=W, image_height=H,
                   tanfovx=np.tan(float(z[f"fovx_{i}"]) / 2), tanfovy=np.tan(float(z[f"fovy_{i}"]) / 2), sh_degree=0)
        # points in front of the camera: p = c2w @ (x,y,z,1)
        pv = np.stack([rng.uniform(-1, 1, 50), rng.uniform(-1, 1, 50), rng.uniform(2, 6, 50), np.ones(50)], 1)
        pw = pv @ np.linalg.inv(V)
        pre = o.preprocess_forward(pw[:, :3], np.full((50, 3), 0.05), np.tile([1.0, 0, 0, 0], (50, 1)),
                                   np.zeros((50, 16, 3)), np.full((50, 1), 0.5), cam)
        hom = pw @ PM
        ndc = hom[:, :2] / (hom[:, 3:4] + 1e-7)
        expect = np.stack([((ndc[:, 0] + 1) * W - 1) / 2, ((ndc[:, 1] + 1) * H - 1) / 2], 1)
        vis = pre["radii"] > 0
        assert vis.sum() > 5
        np.testing.assert_allclose(pre["means2D"][vis], expect[vis], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(pre["depths"][vis], pv[vis, 2], rtol=1e-6, atol=1e-6)


def test_loss_matches_reference_loss_utils():
    z = np.load(os.path.join(G, "loss.npz"))
    o = Oracle(np.float32)
    gt = np.clip(z["gt_u8"].astype(np.float32) / np.float32(255.0), 0, 1)
    l1, ss, grad = o.loss(z["img"], gt, int(z["n_pix"]), 0.2)
    assert abs(l1 - float(z["l1"])) < 1e-6 * abs(float(z["l1"])) + 1e-9
    assert abs(ss - float(z["ssim"])) < 2e-5 * abs(float(z["ssim"]))
    loss = 0.8 * l1 + 0.2 * (1 - ss)
    assert abs(loss - float(z["loss"])) < 1e-6
    np.testing.assert_allclose(grad, z["grad"], rtol=2e-3, atol=2e-9)


def test_covariance_matches_reference_build_scaling_rotation():
    """cov3d_from (oracle/gs_oracle.c) == strip_symmetric(L L^T), L = build_rotation(q) diag(s), executed from the
    reference's own source (utils/general_utils.py:400-451; tests/golden/make_golden.py): pins the (w,x,y,z) convention,
    the element order (xx, xy, xz, yy, yz, zz) and Sigma = R S S^T R^T."""
    z = np.load(os.path.join(G, "cov3d.npz"))
    n = z["scales"].shape[0]
    q = z["rotations"].astype(np.float64)
    q = (q / np.sqrt((q * q).sum(1, keepdims=True)))
    cam = syn.make_camera(320, 200)
    pts = np.tile(np.array([[0.0, 0.0, 5.0]]), (n, 1))          # all in front of the camera: every covariance is written
    for dt in (np.float64, np.float32):
        o = Oracle(dt)
        pre = o.preprocess_forward(pts, z["scales"], q, np.zeros((n, 16, 3)), np.full((n, 1), 0.5), cam)
        assert (pre["radii"] > 0).all()
        # the golden itself is fp32 (off-diagonal elements cancel): the bar is relative to each covariance's largest element
        err = np.abs(pre["cov3D"] - z["cov6"]) / np.abs(z["cov6"]).max(axis=1, keepdims=True)
        assert err.max() < (2e-6 if dt == np.float64 else 4e-6), err.max()
    # the rotation matrix itself, rebuilt from the covariance of unit scales along one axis at a time
    R = z["R"].astype(np.float64)
    for axis in range(3):
        s = np.full((n, 3), 1e-3); s[:, axis] = 1.0
        pre = Oracle(np.float64).preprocess_forward(pts, s, q, np.zeros((n, 16, 3)), np.full((n, 1), 0.5), cam)
        c = pre["cov3D"]
        full = np.stack([c[:, [0, 1, 2]], c[:, [1, 3, 4]], c[:, [2, 4, 5]]], 1)
        col = R[:, :, axis]
        np.testing.assert_allclose(full, col[:, :, None] * col[:, None, :], rtol=0, atol=3e-6)
