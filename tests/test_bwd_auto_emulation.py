"""CPU: a numpy lane-by-lane transcription of the EXPERIMENTAL kernel k_blend_bwd_auto (csrc/blend.cu) -- staging in
batches of 32, per-half-warp candidate ballots from the 4x4-block cull masks, the (T, behind-colour) recurrences with
alpha forced to 0 for skipping lanes, the 16-lane reduction, and the lane roles of the final RED.ADD -- run on a small
scene and compared with the oracle's render_backward.  It validates the kernel's LOGIC (control flow, index maps,
roles) before any device time is spent; the CUDA text itself is validated on the device
(tests/test_gpu_parity.py::test_experimental_backward_wht_parity, GS_B200_EXPERIMENTAL=1)."""
import numpy as np

from gs_b200 import synthetic as syn
from oracle.oracle import Oracle
from test_wht_algebra import LANE, half_reduce9

F = np.float32
ALPHA_MIN, ALPHA_MAX, THR_MARGIN = F(1.0 / 255.0), F(0.99), F(0.02)


def make_records(m2, co, rgb):
    """k_count_tiles: packed record (mx,my,a',b') (c',o,thr,red) (green,blue,ex,ey)."""
    A, B, C, o = (co[:, k].astype(F) for k in range(4))
    thr = (-np.log(F(255.0) * np.maximum(o, F(1e-30))) - THR_MARGIN).astype(F)
    t = (F(-2.0) * thr).astype(np.float64)
    det = A.astype(np.float64) * C.astype(np.float64) - B.astype(np.float64) ** 2
    ok = (t > 0) & (det > 0) & (A > 0) & (C > 0)
    with np.errstate(invalid="ignore", divide="ignore"):
        ex = np.where(t > 0, np.where(ok, np.sqrt(t * C / det) * 1.02 + 0.5, 3.0e38), -1.0).astype(F)
        ey = np.where(t > 0, np.where(ok, np.sqrt(t * A / det) * 1.02 + 0.5, 3.0e38), -1.0).astype(F)
    return dict(mx=m2[:, 0].astype(F), my=m2[:, 1].astype(F), ap=(F(-0.5) * A).astype(F), bp=(-B).astype(F),
                cp=(F(-0.5) * C).astype(F), o=o, thr=thr, r=rgb[:, 0].astype(F), g=rgb[:, 1].astype(F), b=rgb[:, 2].astype(F),
                ex=ex, ey=ey)


def block_mask16(mx, my, ex, ey, X0, Y0):
    if ex < 0:
        return 0
    xl, xh, yl, yh = mx - ex - X0, mx + ex - X0, my - ey - Y0, my + ey - Y0
    xm = sum(1 << b for b in range(4) if xh >= 4.0 * b and xl <= 4.0 * b + 3.0)
    return sum(xm << (4 * b) for b in range(4) if yh >= 4.0 * b and yl <= 4.0 * b + 3.0)


def emulate_tile(tile, gx, W, H, rec, ids, rng, final_T, n_contrib, dimg, bg, out):
    """One CTA of k_blend_bwd_auto: 8 independent warps."""
    X0, Y0 = (tile % gx) * 16, (tile // gx) * 16
    HW = H * W
    ddx, ddy = F(0.5 * W), F(0.5 * H)
    l16, half = LANE & 15, LANE >> 4
    # lane roles: (target array, element, stride, kA, kB, kC, kK, kO, takes_v8)
    roles = {0: ("m", 0, 2 * ddx, 0, 0, 0, 0, False), 1: ("m", 1, 0, ddy, 0, 0, 0, False), 2: ("m", 0, 0, ddx, 0, 0, 0, False),
             3: ("m", 1, 0, 0, 2 * ddy, 0, 0, False), 4: ("c", 0, 0, 0, 0, -0.5, 0, False), 5: ("r", 2, 0, 0, 0, 1.0, 0, True),
             6: ("c", 1, 0, 0, 0, -1.0, 0, False), 8: ("c", 2, 0, 0, 0, -0.5, 0, False), 10: ("c", 3, 0, 0, 0, 0, 1.0, False),
             12: ("r", 0, 0, 0, 0, 1.0, 0, False), 14: ("r", 1, 0, 0, 0, 1.0, 0, False)}
    for warp in range(8):
        by, bx = warp >> 1, (warp & 1) * 2 + half                       # where_am_i
        blk = by * 4 + bx
        px, py = X0 + bx * 4 + (l16 & 3), Y0 + by * 4 + (l16 >> 2)
        inside = (px < W) & (py < H)
        pix = np.where(inside, py * W + px, 0)
        T_final = np.where(inside, final_T.reshape(-1)[pix], 0).astype(F)
        last = np.where(inside, n_contrib.reshape(-1)[pix], 0).astype(np.int64)
        dp = [np.where(inside, dimg.reshape(3, -1)[c][pix], 0).astype(F) for c in range(3)]
        bgdot = (bg[0] * dp[0] + bg[1] * dp[1] + bg[2] * dp[2]).astype(F)
        pxf, pyf = px.astype(F), py.astype(F)
        wlast = int(last.max())
        blkA = int(blk[0])                                              # block of lanes 0-15; lanes 16-31 own blkA + 1
        T, B0, B1, B2 = T_final.copy(), np.zeros(32, F), np.zeros(32, F), np.zeros(32, F)
        g0 = (wlast - 1) & ~31
        while g0 >= 0 and wlast > 0:
            e = g0 + LANE
            valid = e < wlast
            gid = np.where(valid, ids[rng[0] + np.minimum(e, wlast - 1)], 0)
            m16 = np.array([block_mask16(rec["mx"][g], rec["my"][g], rec["ex"][g], rec["ey"][g], X0, Y0) if v else 0
                            for g, v in zip(gid, valid)])
            cA = sum(1 << l for l in range(32) if (m16[l] >> blkA) & 1)
            cB = sum(1 << l for l in range(32) if (m16[l] >> (blkA + 1)) & 1)
            mine = [cA, cB]                                             # per half
            last_rel = last - g0
            while mine[0] or mine[1]:
                has_h = [mine[0] != 0, mine[1] != 0]
                j_h = [mine[h].bit_length() - 1 if has_h[h] else 0 for h in range(2)]
                for h in range(2):
                    if has_h[h]:
                        mine[h] &= ~(1 << j_h[h])
                j = np.where(half == 0, j_h[0], j_h[1])
                has = np.where(half == 0, has_h[0], has_h[1])
                g = gid[j]                                              # stage[j] of this warp (always a valid entry for j = 0)
                a_x, a_y, a_z, a_w = rec["mx"][g], rec["my"][g], rec["ap"][g], rec["bp"][g]
                b_x, b_y, b_z, b_w = rec["cp"][g], rec["o"][g], rec["thr"][g], rec["r"][g]
                dx, dy = (a_x - pxf).astype(F), (a_y - pyf).astype(F)
                power = (dx * (a_z * dx + a_w * dy) + b_x * dy * dy).astype(F)
                ok = has & (j < last_rel) & (power >= b_z)
                if not ok.any():
                    continue
                G = np.exp(power.astype(np.float64)).astype(F)
                alpha = np.minimum(ALPHA_MAX, (b_y * G).astype(F))
                ok = ok & (power <= 0) & (alpha >= ALPHA_MIN)
                if not ok.any():
                    continue
                ae = np.where(ok, alpha, F(0)).astype(F)
                inv = (F(1) / (F(1) - ae)).astype(F)
                T = (T * inv).astype(F)
                d0, d1, d2 = (b_w - B0).astype(F), (rec["g"][g] - B1).astype(F), (rec["b"][g] - B2).astype(F)
                dL_dalpha = ((d0 * dp[0] + d1 * dp[1] + d2 * dp[2]) * T - (T_final * inv) * bgdot).astype(F)
                mw = np.where(ok, b_y * dL_dalpha * G, F(0)).astype(F)
                dch = (ae * T).astype(F)
                B0, B1, B2 = (B0 + ae * d0).astype(F), (B1 + ae * d1).astype(F), (B2 + ae * d2).astype(F)
                mx_, my_ = mw * dx, mw * dy
                v = [mx_, my_, mx_ * dx, mx_ * dy, my_ * dy, mw, dch * dp[0], dch * dp[1], dch * dp[2]]
                v0, v8 = half_reduce9(v)
                for h in range(2):
                    if not ok[16 * h:16 * h + 16].any():                # (okb & half_lanes) == 0: this half adds nothing
                        continue
                    for r16, (arr, elem, kA, kB, kC, kK, kO, use8) in roles.items():
                        lane = 16 * h + r16
                        gg = g[lane]
                        coef = kA * a_z[lane] + kB * a_w[lane] + kC * b_x[lane] + kK + kO / b_y[lane]
                        out[arr][gg, elem] += coef * (v8[lane] if use8 else v0[lane])
            g0 -= 32


def block_mask8(mx, my, ex, ey, X0, Y0):
    """bit w <=> the bbox overlaps warp w's 8x4 block (blend.cu: block_mask)."""
    if ex < 0:
        return 0
    xl, xh, yl, yh = mx - ex - X0, mx + ex - X0, my - ey - Y0, my + ey - Y0
    xm = (1 if (xh >= 0 and xl <= 7) else 0) | (2 if (xh >= 8 and xl <= 15) else 0)
    return sum(xm << (2 * wy) for wy in range(4) if yh >= 4.0 * wy and yl <= 4.0 * wy + 3.0)


def emulate_tile_wht(tile, gx, W, H, rec, ids, rng, final_T, n_contrib, dimg, bg, out):
    """One CTA of k_blend_bwd_wht: per warp an 8x4 block, candidates walked back to front, the weight m through the
    Walsh-Hadamard butterfly, the flush thread rebuilding the moments about the splat centre (chunking is irrelevant
    for the result: a flush sums the warps' partials of one entry)."""
    from test_wht_algebra import flush, warp_partial
    X0, Y0 = (tile % gx) * 16, (tile // gx) * 16
    ddx, ddy = 0.5 * W, 0.5 * H
    n_total = 0
    state = []
    for warp in range(8):
        px, py = X0 + (warp & 1) * 8 + (LANE & 7), Y0 + (warp >> 1) * 4 + (LANE >> 3)     # pixel_of_thread
        inside = (px < W) & (py < H)
        pix = np.where(inside, py * W + px, 0)
        T_final = np.where(inside, final_T.reshape(-1)[pix], 0).astype(F)
        last = np.where(inside, n_contrib.reshape(-1)[pix], 0).astype(np.int64)
        dp = [np.where(inside, dimg.reshape(3, -1)[c][pix], 0).astype(F) for c in range(3)]
        state.append(dict(px=px.astype(F), py=py.astype(F), T_final=T_final, last=last, dp=dp,
                          bgdot=(bg[0] * dp[0] + bg[1] * dp[1] + bg[2] * dp[2]).astype(F), T=T_final.copy(),
                          B=[np.zeros(32, F) for _ in range(3)]))
        n_total = max(n_total, int(last.max()))
    for e in range(n_total - 1, -1, -1):
        g = int(ids[rng[0] + e])
        mask = block_mask8(rec["mx"][g], rec["my"][g], rec["ex"][g], rec["ey"][g], X0, Y0)
        s = np.zeros(9)
        any_partial = False
        for warp in range(8):
            st = state[warp]
            if not (mask >> warp) & 1 or e >= int(st["last"].max()):
                continue
            dx, dy = (rec["mx"][g] - st["px"]).astype(F), (rec["my"][g] - st["py"]).astype(F)
            power = (dx * (rec["ap"][g] * dx + rec["bp"][g] * dy) + rec["cp"][g] * dy * dy).astype(F)
            ok = (e < st["last"]) & (power >= rec["thr"][g])
            if not ok.any():
                continue
            G = np.exp(power.astype(np.float64)).astype(F)
            alpha = np.minimum(ALPHA_MAX, (rec["o"][g] * G).astype(F))
            ok = ok & (power <= 0) & (alpha >= ALPHA_MIN)
            if not ok.any():
                continue
            ae = np.where(ok, alpha, F(0)).astype(F)
            inv = (F(1) / (F(1) - ae)).astype(F)
            st["T"] = (st["T"] * inv).astype(F)
            d = [(rec[c][g] - st["B"][k]).astype(F) for k, c in enumerate("rgb")]
            dL_dalpha = ((d[0] * st["dp"][0] + d[1] * st["dp"][1] + d[2] * st["dp"][2]) * st["T"] -
                         (st["T_final"] * inv) * st["bgdot"]).astype(F)
            mw = np.where(ok, rec["o"][g] * dL_dalpha * G, F(0)).astype(F)
            dch = (ae * st["T"]).astype(F)
            for k in range(3):
                st["B"][k] = (st["B"][k] + ae * d[k]).astype(F)
            cf = warp_partial(mw, [dch * st["dp"][0], dch * st["dp"][1], dch * st["dp"][2]], np.float32)
            ux, uy = rec["mx"][g] - F(X0 + (warp & 1) * 8), rec["my"][g] - F(Y0 + (warp >> 1) * 4)
            s[:6] += flush(cf, ux, uy, np.float32)
            s[6:] += cf[16:19]
            any_partial = True
        if any_partial:
            ap, bp, cp, o = rec["ap"][g], rec["bp"][g], rec["cp"][g], rec["o"][g]
            out["m"][g, 0] += (2 * ap * s[0] + bp * s[1]) * ddx
            out["m"][g, 1] += (2 * cp * s[1] + bp * s[0]) * ddy
            out["c"][g] += (-0.5 * s[2], -s[3], -0.5 * s[4], s[5] / o)
            out["r"][g] += s[6:9]


def test_wht_backward_logic_matches_the_oracle():
    _run(emulate_tile_wht)


def test_autonomous_backward_logic_matches_the_oracle():
    _run(emulate_tile)


def _run(emulate):
    W, H, n = 64, 48, 1500
    cam = syn.make_camera(W, H, yaw_deg=2.0)
    sc = syn.make_scene(n, W, H, seed=8, radius_px=6.0)
    o = Oracle(np.float32)
    pre = o.preprocess_forward(sc["means3D"], sc["scales"], sc["rotations"], sc["shs"], sc["opacities"], cam)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    bg = np.array([0.2, 0.4, 0.1], F)
    fwd = o.render_forward(H, W, pre["means2D"], pre["conic_opacity"], pre["rgb"], pre["depths"], pre["radii"],
                           np.ones(gx * gy, np.uint8), tuple(bg))
    dimg = np.random.default_rng(5).normal(size=(3, H, W)).astype(F)
    ref = o.render_backward(H, W, pre["means2D"], pre["conic_opacity"], pre["rgb"], tuple(bg), fwd, dimg)
    rec = make_records(pre["means2D"], pre["conic_opacity"], pre["rgb"])
    out = {"m": np.zeros((n, 2)), "c": np.zeros((n, 4)), "r": np.zeros((n, 3))}
    ids = fwd["ids"].astype(np.int64)
    ranges = fwd["ranges"].reshape(-1, 2)
    assert fwd["R"] > 2000
    for tile in range(gx * gy):
        emulate(tile, gx, W, H, rec, ids, ranges[tile], fwd["final_T"], fwd["n_contrib"], dimg, bg, out)
    for key, name in (("m", "means2D"), ("c", "conic_opacity"), ("r", "rgb")):
        a, b = out[key], ref[name].astype(np.float64)
        scale = np.sqrt((b ** 2).mean())
        bad = np.abs(a - b) > 2e-4 * np.abs(b) + 2e-4 * scale
        assert bad.mean() <= 2e-3, (name, float(bad.mean()), float(np.abs(a - b).max()), float(scale))
