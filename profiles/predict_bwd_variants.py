"""Predicts the issued-instruction totals of the backward blend kernels (default and experimental variants) on a scaled
copy of the c2 workload, from (a) the per-iteration instruction counts read off the SASS (profiles/r1_sass_blend_bwd.md)
and (b) the exact iteration counts each kernel's control flow produces on the scene (which candidates every warp /
half-warp walks and where each iteration exits), computed here on the CPU with the oracle's forward.

The blend kernels are issue bound (ncu: issue slots ~80 % busy), so relative instruction totals are a first-order
prediction of relative run time -- barriers, atomics and occupancy are NOT modelled; tests/ab_experimental.py measures
the real thing.  Analysis tool (test infrastructure: it imports oracle/), not product code.

    python profiles/predict_bwd_variants.py            # ~1 minute
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "grendel-gs_b200"), os.path.join(ROOT, "tests")]
from gs_b200 import synthetic as syn  # noqa: E402
from oracle.oracle import Oracle      # noqa: E402
from test_bwd_auto_emulation import make_records  # noqa: E402

F = np.float32
# issued instructions per loop iteration by exit point (first vote fails / second vote fails / fully processed), SASS
COST = {"default": (22, 30, 111), "wht64": (22, 30, 98), "wht128": (22, 30, 92), "auto": (27, 40, 136)}
GROUP_OVERHEAD = 12        # per (warp, 32-entry group): ballot of the cull masks, loop bookkeeping
STAGE_CTA = 8 * 25         # default / wht: 256 threads stage <= CHUNK records (warp-instructions per chunk)
STAGE_WARP = 45            # auto: one warp stages 32 records
FLUSH_FIXED, FLUSH_DEFAULT, FLUSH_WHT = 64, 18, 100   # per flush thread: mask tests + atomics; per warp partial


def masks(rec, g, X0, Y0):
    """(n,) cull masks of the entries for 8x4 blocks (8 bits) and 4x4 blocks (16 bits)."""
    mx, my, ex, ey = rec["mx"][g], rec["my"][g], rec["ex"][g], rec["ey"][g]
    xl, xh, yl, yh = mx - ex - X0, mx + ex - X0, my - ey - Y0, my + ey - Y0
    live = ex >= 0
    m8 = np.zeros(len(g), np.int64)
    m16 = np.zeros(len(g), np.int64)
    for wy in range(4):
        rowhit = (yh >= 4.0 * wy) & (yl <= 4.0 * wy + 3.0) & live
        for wx in range(2):
            m8 |= (rowhit & (xh >= 8.0 * wx) & (xl <= 8.0 * wx + 7.0)).astype(np.int64) << (2 * wy + wx)
        for bx in range(4):
            m16 |= (rowhit & (xh >= 4.0 * bx) & (xl <= 4.0 * bx + 3.0)).astype(np.int64) << (4 * wy + bx)
    return m8, m16


def main():
    W, H, n = 480, 270, 125_000     # c2 (1920x1080, 2 M Gaussians) scaled by 1/16 in pixels and Gaussians: same per-tile lists
    cam = syn.make_camera(W, H)
    sc = syn.make_scene(n, W, H, seed=0)
    o = Oracle(np.float32, threads=max(1, (os.cpu_count() or 8) // 2))
    pre = o.preprocess_forward(sc["means3D"], sc["scales"], sc["rotations"], sc["shs"], sc["opacities"], cam)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    fwd = o.render_forward(H, W, pre["means2D"], pre["conic_opacity"], pre["rgb"], pre["depths"], pre["radii"],
                           np.ones(gx * gy, np.uint8), (0, 0, 0))
    rec = make_records(pre["means2D"], pre["conic_opacity"], pre["rgb"])
    ids, ranges = fwd["ids"].astype(np.int64), fwd["ranges"].reshape(-1, 2)
    print(f"scene: {n} Gaussians @ {W}x{H}: R = {fwd['R']} instances, {fwd['R'] / (gx * gy):.0f} per tile "
          f"(c2 on the device: 5.74 M / 8160 = 703)")
    tot = {k: 0.0 for k in COST}
    iters = {k: np.zeros(3) for k in COST}
    lanes_useful = {"8x4": [0, 0], "4x4": [0, 0]}
    yy, xx = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    for tile in range(gx * gy):
        X0, Y0 = (tile % gx) * 16, (tile // gx) * 16
        beg, end = ranges[tile]
        if end <= beg:
            continue
        g = ids[beg:end]
        px, py = (X0 + xx).reshape(-1), (Y0 + yy).reshape(-1)
        inside = (px < W) & (py < H)
        last = np.where(inside, fwd["n_contrib"][np.minimum(py, H - 1), np.minimum(px, W - 1)], 0).astype(np.int64)
        n_total = int(last.max())
        if n_total == 0:
            continue
        g = g[:n_total]
        e = np.arange(n_total)
        dx = rec["mx"][g][None, :] - px[:, None].astype(F)
        dy = rec["my"][g][None, :] - py[:, None].astype(F)
        power = dx * (rec["ap"][g] * dx + rec["bp"][g] * dy) + rec["cp"][g] * dy * dy
        ok1 = (e[None, :] < last[:, None]) & (power >= rec["thr"][g])
        alpha = np.minimum(0.99, rec["o"][g] * np.exp(np.minimum(power, 0.0)))
        ok2 = ok1 & (power <= 0) & (alpha >= 1.0 / 255.0)
        m8, m16 = masks(rec, g, X0, Y0)
        lx, ly = px - X0, py - Y0
        # ---- default / wht: one 8x4 block per warp, CHUNK-entry chunks ---------------------------------------------
        for name, chunk in (("default", 128), ("wht64", 64), ("wht128", 128)):
            c1, c2, c3 = COST[name]
            n_chunks = (n_total + chunk - 1) // chunk
            cost = n_chunks * STAGE_CTA
            partial = np.zeros(n_total, np.int64)
            for w in range(8):
                sel = ((lx // 8) == (w & 1)) & ((ly // 4) == (w >> 1))
                wl = int(last[sel].max())
                if wl == 0:
                    continue
                reach = min(n_total, ((wl + chunk - 1) // chunk) * chunk)   # chunks with base < wlast are walked
                cand = ((m8[:reach] >> w) & 1) == 1
                a1 = ok1[sel][:, :reach].any(axis=0) & cand
                a2 = ok2[sel][:, :reach].any(axis=0) & cand
                k3, k2, k1 = int(a2.sum()), int((a1 & ~a2).sum()), int((cand & ~a1).sum())
                cost += k1 * c1 + k2 * c2 + k3 * c3 + ((reach + 31) // 32) * GROUP_OVERHEAD
                iters[name] += (k1, k2, k3)
                partial[:reach] += a2
                if name == "default":
                    lanes_useful["8x4"][0] += int(ok2[sel][:, :reach][:, a2].sum())
                    lanes_useful["8x4"][1] += 32 * k3
            per_partial = FLUSH_DEFAULT if name == "default" else FLUSH_WHT
            for c in range(n_chunks):
                pc = partial[c * chunk:(c + 1) * chunk]
                for q in range(0, len(pc), 32):                 # one flush warp per 32 entries: pays for its busiest entry
                    cost += FLUSH_FIXED + per_partial * int(pc[q:q + 32].max())
            tot[name] += cost
        # ---- auto: two 4x4 blocks per warp, 32-entry batches, candidates of the two halves paired in order ---------------
        c1, c2, c3 = COST["auto"]
        cost = 0
        for w in range(8):
            blkA = (w >> 1) * 4 + (w & 1) * 2
            selA = ((ly // 4) == (w >> 1)) & ((lx // 4) == (w & 1) * 2)
            selB = ((ly // 4) == (w >> 1)) & ((lx // 4) == (w & 1) * 2 + 1)
            wl = int(max(last[selA].max(), last[selB].max()))
            for g0 in range(0, wl, 32):
                hi = min(g0 + 32, wl)
                cost += STAGE_WARP
                seq = []
                for sel, blk in ((selA, blkA), (selB, blkA + 1)):
                    idx = np.nonzero(((m16[g0:hi] >> blk) & 1) == 1)[0][::-1] + g0       # back to front
                    seq.append((idx, ok1[sel][:, idx].any(axis=0), ok2[sel][:, idx]))
                nit = max(len(seq[0][0]), len(seq[1][0]))
                for k in range(nit):
                    s1 = any(k < len(s[0]) and s[1][k] for s in seq)
                    s2 = any(k < len(s[0]) and s[2][:, k].any() for s in seq)
                    cost += c3 if s2 else (c2 if s1 else c1)
                    iters["auto"][2 if s2 else (1 if s1 else 0)] += 1
                    if s2:
                        lanes_useful["4x4"][0] += sum(int(s[2][:, k].sum()) for s in seq if k < len(s[0]))
                        lanes_useful["4x4"][1] += 32
        tot["auto"] += cost
    base = tot["default"]
    print("variant    warp-instructions   vs default   iterations (exit 1 / exit 2 / full)")
    for k in COST:
        print(f"{k:9s}  {tot[k]:16.3e}   {tot[k] / base:9.3f}   {iters[k].astype(int).tolist()}")
    for k, (u, t) in lanes_useful.items():
        print(f"useful lanes in fully processed iterations, {k} blocks: {u / max(t, 1):.2f}")


if __name__ == "__main__":
    main()
