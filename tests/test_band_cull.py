"""Host-side check of the blend kernels' per-block culling (csrc/blend.cu: ellipse_bands / block_mask16 / block_mask):
an fp32 numpy transcription of the row-band extents, run over whole tile lists of ordinary, anisotropic and needle-shaped
scenes, must never drop a (block, splat) pair in which some pixel passes the kernels' exponent test
(`thr <= power <= 0`, evaluated here in fp32 without FMA and in fp64) -- and it has to remove a useful share of the
bounding-box candidates, or it is not worth its instructions.  The device-side proof (culled == unculled, bit for bit) is
tests/test_gpu_parity.py::test_needle_splats_survive_block_culling and ::test_block_cull_is_invisible."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "grendel-gs_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from gs_b200 import synthetic as syn  # noqa: E402
from oracle.oracle import Oracle      # noqa: E402

F = np.float32
THR_MARGIN = F(0.0)    # k_count_tiles: thr = ln(1/(255 opacity)), the exponent test IS the alpha >= 1/255 test
ABS_MARGIN = F(0.05)


def make_records(m2, co):
    """k_count_tiles' record fields (csrc/binning.cu), fp32, with the compensated determinant in fp64."""
    A, B, C, o = (co[:, k].astype(F) for k in range(4))
    thr = (-np.log(F(255.0) * np.maximum(o, F(1e-30))) - THR_MARGIN).astype(F)
    t = (F(-2.0) * thr).astype(np.float64)
    det = A.astype(np.float64) * C.astype(np.float64) - B.astype(np.float64) ** 2
    ok = (t > 0) & (det > 0) & (A > 0) & (C > 0)
    with np.errstate(invalid="ignore", divide="ignore"):
        ex = np.where(t > 0, np.where(ok, np.sqrt(t * C / det) * 1.02 + 0.5, 3.0e38), -1.0).astype(F)
        ey = np.where(t > 0, np.where(ok, np.sqrt(t * A / det) * 1.02 + 0.5, 3.0e38), -1.0).astype(F)
    return dict(mx=m2[:, 0].astype(F), my=m2[:, 1].astype(F), az=(F(-0.5) * A).astype(F), aw=(-B).astype(F),
                bx=(F(-0.5) * C).astype(F), thr=thr, ex=ex, ey=ey)


def ellipse_bands(rec, g, X0, Y0):
    """fp32 transcription of ellipse_bands(): [xl, xh] per band of four pixel rows, tile coordinates."""
    mx, my = rec["mx"][g], rec["my"][g]
    az, aw, bx, thr, ex, ey = (rec[k][g] for k in ("az", "aw", "bx", "thr", "ex", "ey"))
    out = []
    with np.errstate(all="ignore"):
        ex0, ey0 = ((ex - F(0.5)) * F(1 / 1.02)).astype(F), ((ey - F(0.5)) * F(1 / 1.02)).astype(F)
        mgx, mgy = (F(0.02) * ex0 + ABS_MARGIN).astype(F), (F(0.02) * ey0 + ABS_MARGIN).astype(F)
        inv_a = (F(1) / -az).astype(F)
        kappa = (F(0.5) * aw * inv_a).astype(F)
        hw2 = (-thr * inv_a).astype(F)
        v_r = (F(-0.5) * aw * ex0 * (F(1) / bx)).astype(F)
        inv_ey0 = (F(1) / ey0).astype(F)
        cx = (mx - F(X0)).astype(F)
        exact = ex < F(1e30)
        for q in range(4):
            v0 = (F(Y0 + 4 * q) - my).astype(F)
            v1 = (v0 + F(3)).astype(F)
            inband = (ex >= 0) & (v1 >= -ey) & (v0 <= ey)
            lo, hi = (v0 - mgy).astype(F), (v1 + mgy).astype(F)
            vh = np.fmin(np.fmax(v_r, lo), hi)
            vl = np.fmin(np.fmax(-v_r, lo), hi)
            qh, ql = (vh * inv_ey0).astype(F), (vl * inv_ey0).astype(F)
            sh = np.fmax(F(0), (F(1) - qh * qh).astype(F))
            sl = np.fmax(F(0), (F(1) - ql * ql).astype(F))
            h = (cx + (kappa * vh + np.sqrt(hw2 * sh) + mgx)).astype(F)
            l = (cx + (kappa * vl - np.sqrt(hw2 * sl) - mgx)).astype(F)
            h, l = np.where(exact, h, F(1e30)), np.where(exact, l, F(-1e30))
            out.append((np.where(inband, l, F(1e30)), np.where(inband, h, F(-1e30))))
    return out


def needle_scene(n, W, H):
    rng = np.random.default_rng(42)
    sc = syn.make_scene(n, W, H, seed=9, radius_px=6.0)
    sc["means3D"][:, 2] = rng.uniform(3.0, 6.0, n)
    sc["means3D"][:, 0] = rng.uniform(-1.0, 1.0, n)
    sc["means3D"][:, 1] = rng.uniform(-0.6, 0.6, n)
    long_axis = rng.uniform(2.0, 40.0, n)
    sc["scales"] = np.stack([long_axis, np.full(n, 2e-4), np.full(n, 2e-4)], 1).astype(np.float32)
    ang = rng.uniform(0, np.pi, n)
    sc["rotations"] = np.stack([np.cos(ang / 2), np.zeros(n), np.zeros(n), np.sin(ang / 2)], 1).astype(np.float32)
    sc["opacities"] = rng.uniform(0.05, 0.9, (n, 1)).astype(np.float32)
    return sc


def run_scene(sc, W, H):
    cam = syn.make_camera(W, H)
    o = Oracle(np.float32, threads=min(8, os.cpu_count() or 1))
    pre = o.preprocess_forward(sc["means3D"], sc["scales"], sc["rotations"], sc["shs"], sc["opacities"], cam)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    fwd = o.render_forward(H, W, pre["means2D"], pre["conic_opacity"], pre["rgb"], pre["depths"], pre["radii"],
                           np.ones(gx * gy, np.uint8), (0, 0, 0))
    rec = make_records(pre["means2D"], pre["conic_opacity"])
    ids, ranges = fwd["ids"].astype(np.int64), fwd["ranges"].reshape(-1, 2)
    yy, xx = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    lx, ly = xx.reshape(-1), yy.reshape(-1)
    tot = dict(box=0, kept=0, needed=0, dropped_needed=0)
    for tile in range(gx * gy):
        beg, end = ranges[tile]
        if end <= beg:
            continue
        X0, Y0 = (tile % gx) * 16, (tile // gx) * 16
        g = ids[beg:end]
        px, py = (X0 + lx), (Y0 + ly)
        need = np.zeros((256, len(g)), bool)
        for dt in (F, np.float64):   # the kernels' test under two roundings of the same fp32 inputs
            dx = rec["mx"][g][None, :].astype(dt) - px[:, None].astype(dt)
            dy = rec["my"][g][None, :].astype(dt) - py[:, None].astype(dt)
            power = dx * (rec["az"][g].astype(dt) * dx + rec["aw"][g].astype(dt) * dy) + rec["bx"][g].astype(dt) * dy * dy
            need |= (power >= rec["thr"][g]) & (power <= 0)
        mx, my, ex, ey = rec["mx"][g], rec["my"][g], rec["ex"][g], rec["ey"][g]
        xl, xh, yl, yh = mx - ex - X0, mx + ex - X0, my - ey - Y0, my + ey - Y0
        bands = ellipse_bands(rec, g, X0, Y0)
        for q in range(4):
            bl, bh = bands[q]
            for bw in (4, 8):
                for bx in range(16 // bw):
                    box = (ex >= 0) & (yh >= 4 * q) & (yl <= 4 * q + 3) & (xh >= bw * bx) & (xl <= bw * bx + bw - 1)
                    keep = (bh >= bw * bx) & (bl <= bw * bx + bw - 1)
                    sel = ((lx // bw) == bx) & ((ly // 4) == q)
                    needed = need[sel].any(axis=0)
                    tot["box"] += int(box.sum())
                    tot["kept"] += int(keep.sum())
                    tot["needed"] += int(needed.sum())
                    tot["dropped_needed"] += int((needed & ~keep).sum())
    return tot


@pytest.mark.parametrize("name", ["ordinary", "large", "anisotropic", "needles"])
def test_band_cull_is_conservative(name):
    if name == "ordinary":
        W, H = 240, 144
        sc = syn.make_scene(20_000, W, H, seed=0)
    elif name == "large":
        W, H = 160, 96
        sc = syn.make_scene(3_000, W, H, seed=3, radius_px=25.0)
    elif name == "anisotropic":
        W, H = 240, 144
        sc = syn.make_scene(15_000, W, H, seed=5, radius_px=5.0)
        sc["scales"][:, 0] *= 8.0
        sc["scales"][:, 1] /= 8.0
    else:
        W, H = 256, 160
        sc = needle_scene(400, W, H)
    t = run_scene(sc, W, H)
    print(f"[band cull] {name}: box candidates {t['box']}, kept {t['kept']} ({t['kept'] / max(t['box'], 1):.2f}), "
          f"needed {t['needed']}, needed-but-dropped {t['dropped_needed']}")
    assert t["dropped_needed"] == 0
    assert t["kept"] <= t["box"]
    if name in ("ordinary", "anisotropic"):
        assert t["kept"] < 0.9 * t["box"]   # it has to pay for itself


def test_unsigned_compare_is_the_range_test():
    """The blend kernels test `thr <= power <= 0` as ONE unsigned compare of bit patterns on q = -power, t = -thr >= 0
    (csrc/blend.cu, staging of k_blend_fwd2 / k_blend_bwd_seg): for every fp32 q -- positive, negative, +-0, subnormal, inf,
    NaN of either sign -- `bits(q) <= bits(t)` must equal `0 <= q <= t`, except q = -0.0, which the kernels cannot produce
    (the non-negative gamma' dy^2 is added last: x + (+0) is never -0)."""
    rng = np.random.default_rng(0)
    special = np.array([0.0, np.finfo(F).tiny, 1e-45, 1.0, 5.54, 88.0, 3e38, np.inf, np.nan, -np.nan,
                        -1e-45, -np.finfo(F).tiny, -1.0, -3e38, -np.inf], dtype=F)
    q = np.concatenate([special, rng.normal(0, 10, 20000).astype(F), rng.uniform(0, 12, 20000).astype(F),
                        rng.integers(0, 2**32, 20000, dtype=np.uint64).astype(np.uint32).view(F)])
    q = q[q.view(np.uint32) != 0x80000000]          # -0.0: excluded by construction in the kernels
    for t in (F(0.0), F(1e-3), F(0.5), F(5.5412), F(87.3)):
        with np.errstate(invalid="ignore"):
            want = (q >= 0) & (q <= t)
        got = q.view(np.uint32) <= np.array(t, dtype=F).view(np.uint32)
        assert np.array_equal(want, got), t
    # and the sum that produces q never yields -0: (anything) + (+0 or positive) under round-to-nearest
    a = np.array([-0.0, 0.0, -1e-45, 1e-45], dtype=F)
    for b in (F(0.0), F(1e-45), F(2.0)):
        assert not np.any((a + b).view(np.uint32) == 0x80000000)
