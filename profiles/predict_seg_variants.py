"""Round 2: issued-instruction prediction of the SEGMENT-parallel backward blend designs on a 1/16-scale copy of c2.

Design family: the forward stores a per-pixel checkpoint every K entries of a tile list, so every (tile, K-entry
segment) can be walked back to front independently.  A group of G lanes (G = 32 / 16 / 8) owns one (tile, segment):
lane l holds ONE pixel of each of the 256 / G pixel blocks of the tile, loops over the blocks the splat can reach
(block mask from the {alpha >= 1/255} bounding box, as today), accumulates the nine gradient sums in registers ACROSS
the blocks, reduces ONCE per splat over the G lanes and issues one RED set -- no CTA barrier, no shared partials, no
flush.  32 / G groups share a warp and run in lockstep (an iteration lasts as long as the longest group's).

Cost model: warp instructions per block iteration by exit point (no lane passes the exponent test / no lane blends /
fully processed) + per-splat overhead (record load, reduction, RED); compared with the shipped kernel's model from
profiles/predict_bwd_variants.py (round 1).  Analysis tool: imports oracle/, not product code.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "grendel-gs_b200"), os.path.join(ROOT, "tests")]
from gs_b200 import synthetic as syn  # noqa: E402
from oracle.oracle import Oracle      # noqa: E402

F = np.float32
THR_MARGIN = F(0.02)
BLOCK_COST = (24, 32, 54)                 # exit 1 / exit 2 / full, per block iteration
SPLAT_OVERHEAD = {32: 80, 16: 74, 8: 62}  # record load + zero + premix + G-lane butterfly of 9 values + RED
EMPTY_SPLAT = 10                          # every group's mask is empty: pop the entry, nothing else
BLOCK_SHAPE = {32: (8, 4), 16: (4, 4), 8: (4, 2)}


def make_records(m2, co, rgb):
    A, B, C, o = (co[:, k].astype(F) for k in range(4))
    thr = (-np.log(F(255.0) * np.maximum(o, F(1e-30))) - THR_MARGIN).astype(F)
    t = (F(-2.0) * thr).astype(np.float64)
    det = A.astype(np.float64) * C.astype(np.float64) - B.astype(np.float64) ** 2
    ok = (t > 0) & (det > 0) & (A > 0) & (C > 0)
    with np.errstate(invalid="ignore", divide="ignore"):
        ex = np.where(t > 0, np.where(ok, np.sqrt(t * C / det) * 1.02 + 0.5, 3.0e38), -1.0).astype(F)
        ey = np.where(t > 0, np.where(ok, np.sqrt(t * A / det) * 1.02 + 0.5, 3.0e38), -1.0).astype(F)
    return dict(mx=m2[:, 0].astype(F), my=m2[:, 1].astype(F), ap=(F(-0.5) * A).astype(F), bp=(-B).astype(F),
                cp=(F(-0.5) * C).astype(F), o=o, thr=thr, ex=ex, ey=ey)


def main():
    W, H, n = 480, 270, 125_000
    cam = syn.make_camera(W, H)
    sc = syn.make_scene(n, W, H, seed=0)
    o = Oracle(np.float32, threads=max(1, (os.cpu_count() or 8) // 2))
    pre = o.preprocess_forward(sc["means3D"], sc["scales"], sc["rotations"], sc["shs"], sc["opacities"], cam)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    fwd = o.render_forward(H, W, pre["means2D"], pre["conic_opacity"], pre["rgb"], pre["depths"], pre["radii"],
                           np.ones(gx * gy, np.uint8), (0, 0, 0))
    rec = make_records(pre["means2D"], pre["conic_opacity"], pre["rgb"])
    ids, ranges = fwd["ids"].astype(np.int64), fwd["ranges"].reshape(-1, 2)
    R = int(fwd["R"])
    print(f"scene: {n} Gaussians @ {W}x{H}: R = {R} instances, {R / (gx * gy):.0f} per tile")
    yy, xx = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    results = {}
    for G in (32, 16, 8):
        for K in (32, 64):
            results[(G, K)] = dict(cost=0.0, it=np.zeros(3), splats=0, empty=0, units=0, useful=0, lanes=0)
    walked = 0
    for tile in range(gx * gy):
        X0, Y0 = (tile % gx) * 16, (tile // gx) * 16
        beg, end = ranges[tile]
        if end <= beg:
            continue
        px, py = (X0 + xx).reshape(-1), (Y0 + yy).reshape(-1)
        inside = (px < W) & (py < H)
        last = np.where(inside, fwd["n_contrib"][np.minimum(py, H - 1), np.minimum(px, W - 1)], 0).astype(np.int64)
        n_total = int(last.max())
        if n_total == 0:
            continue
        walked += n_total
        g = ids[beg:end][:n_total]
        e = np.arange(n_total)
        dx = rec["mx"][g][None, :] - px[:, None].astype(F)
        dy = rec["my"][g][None, :] - py[:, None].astype(F)
        power = dx * (rec["ap"][g] * dx + rec["bp"][g] * dy) + rec["cp"][g] * dy * dy
        ok1 = (e[None, :] < last[:, None]) & (power >= rec["thr"][g])
        alpha = np.minimum(0.99, rec["o"][g] * np.exp(np.minimum(power, 0.0)))
        ok2 = ok1 & (power <= 0) & (alpha >= 1.0 / 255.0)
        lx, ly = px - X0, py - Y0
        mx, my, ex, ey = rec["mx"][g], rec["my"][g], rec["ex"][g], rec["ey"][g]
        xl, xh, yl, yh = mx - ex - X0, mx + ex - X0, my - ey - Y0, my + ey - Y0
        for G in (32, 16, 8):
            bw, bh = BLOCK_SHAPE[G]
            nbx, nby = 16 // bw, 16 // bh
            nb = nbx * nby
            # per (block, entry): candidate by bbox; exit class 0/1/2 ; -1 = not a candidate
            cls = np.full((nb, n_total), -1, np.int64)
            for by in range(nby):
                for bx in range(nbx):
                    b = by * nbx + bx
                    cand = (ex >= 0) & (yh >= bh * by) & (yl <= bh * by + bh - 1) & (xh >= bw * bx) & (xl <= bw * bx + bw - 1)
                    sel = ((lx // bw) == bx) & ((ly // bh) == by)
                    blast = int(last[sel].max())
                    cand &= e < blast          # the block's deepest contributor bounds its walk (kept per block in smem)
                    a1 = ok1[sel].any(axis=0)
                    a2 = ok2[sel].any(axis=0)
                    cls[b] = np.where(cand, np.where(a2, 2, np.where(a1, 1, 0)), -1)
                    if G == 16:
                        pass
            cnt = (cls >= 0).sum(axis=0)                         # candidate blocks per entry
            # sorted exit classes per entry, in block order (the kernel walks set bits in order)
            groups = 32 // G
            for K in (32, 64):
                res = results[(G, K)]
                nseg = (n_total + K - 1) // K
                res["units"] += (nseg + groups - 1) // groups
                for u in range(0, nseg, groups):
                    segs = [s for s in range(u, min(u + groups, nseg))]
                    # lockstep: step t of the unit handles entry (seg_end - 1 - t) of every group's segment
                    for t in range(K):
                        ents = [min((s + 1) * K, n_total) - 1 - t for s in segs]
                        ents = [en for s, en in zip(segs, ents) if en >= s * K]
                        if not ents:
                            break
                        res["splats"] += len(ents)
                        cs = [cls[:, en][cls[:, en] >= 0] for en in ents]
                        nit = max(len(c) for c in cs)
                        if nit == 0:
                            res["cost"] += EMPTY_SPLAT
                            res["empty"] += 1
                            continue
                        for k in range(nit):
                            c = max(int(ci[k]) for ci in cs if k < len(ci))
                            res["cost"] += BLOCK_COST[c]
                            res["it"][c] += 1
                        res["cost"] += SPLAT_OVERHEAD[G]
    print(f"entries walked (sum over tiles of the deepest contributor): {walked} = {walked / R:.2f} of R")
    print("G lanes/splat  K   warp-instr (x16 = c2)   per walked entry   iterations exit1/exit2/full   empty splat steps")
    for (G, K), r in results.items():
        print(f"{G:3d} {K:4d}   {r['cost']:.3e} ({16 * r['cost']:.3e})   {r['cost'] / walked:7.1f}   "
              f"{r['it'].astype(int).tolist()}   {r['empty']}")
    print("round-1 model of the shipped kernel on the same scene: 8.64e7 (c2: 1.38e9)")


if __name__ == "__main__":
    main()
