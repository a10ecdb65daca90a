"""Multi-GPU parity script (run under torch.distributed.run, one rank per GPU):
pixel-sharded render + sparse all-to-all + mirrored backward must reproduce what a single rank computes.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mgpu_parity.py

Checks (rank 0 prints PASS / raises):
  1. strips rendered by the W ranks, summed with all_reduce (train_internal.py:466-469), equal the oracle's full
     render; tile lists are partition independent so pixels agree to fp32 rounding, non-local tiles are exactly 0;
  2. the per-shard parameter gradients, gathered, match the oracle evaluated with the SAME strip-wise loss
     (zero-padded SSIM at strip edges, loss_distribution.py:2553-2576) within 1e-4.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "grendel-gs_b200"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

from gs_b200 import division, ops, pipeline, synthetic as syn  # noqa: E402


def main():
    sys.stdout.reconfigure(line_buffering=True)   # progress survives a timeout kill when stdout is a pipe
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    W, H, N, B = 320, 208, 30000, world
    scene = syn.make_scene(N, W, H, seed=21, radius_px=8.0)
    cams = syn.make_batch_cameras(W, H, B)
    gts = [syn.make_gt_image(W, H, seed=50 + k) for k in range(B)]
    tr = pipeline.Trainer(scene, cams, [torch.from_numpy(g).pin_memory() for g in gts], dev, rank, world)
    loss = tr.step(resident=False)

    # ---- 1. image assembly -------------------------------------------------------------------------
    with torch.no_grad():
        p = tr.params
        imgs = []
        strategies, _ = division.start_strategy([c.uid for c in tr.dcams], tr.history, world, rank)
        settings = [c.settings(3) for c in tr.dcams]
        screen = [ops.preprocess_gaussians_raw(p._xyz, p._features_dc, p._features_rest, p._scaling, p._rotation,
                                               p._opacity, rs) for rs in settings]
        stacked = tuple(torch.stack([s_[q] for s_ in screen]) for q in range(5))
        red, _ = tr._ex.exchange(*stacked, strategies, settings, world, rank)
        for k, st in enumerate(strategies):
            img = torch.zeros((3, H, W), device=dev)
            if st.local_rows() is not None:
                m2, rgb, co, radii, depths = red[k]
                img, *_ = ops.render_gaussians(m2, co, rgb, depths, radii, st.get_compute_locally(tr.tile_x, dev), settings[k])
                y0, y1 = st.local_pixel_rows(H)
                outside = torch.ones((H, W), dtype=torch.bool, device=dev)
                outside[y0:y1] = False
                assert (img[:, outside] == 0).all(), "non-local tiles must be exactly zero"
            dist.all_reduce(img)
            imgs.append(img.cpu().numpy())

    # ---- 1b. NVLink peer-memory exchange == all_to_all_single, forward and backward, bit for bit ------------------
    if tr._peer is not None:
        outs = []
        for peer in (tr._peer, None):
            leaves = [stacked[q].detach().clone().requires_grad_(True) for q in range(3)]
            (m2, c3, co, rad, dep), vs, _ = tr._ex.exchange_cat(leaves[0], leaves[1], leaves[2], stacked[3], stacked[4],
                                                                strategies, settings, world, rank, None, peer)
            gen = torch.Generator(device=dev).manual_seed(7 + rank)
            up = [torch.randn(t.shape, device=dev, generator=gen) for t in (m2, c3, co)]
            ((m2 * up[0]).sum() + (c3 * up[1]).sum() + (co * up[2]).sum()).backward()
            outs.append(([m2.detach(), c3.detach(), co.detach(), rad, dep], [l.grad for l in leaves], vs))
        assert outs[0][2] == outs[1][2]
        for a, b in zip(outs[0][0] + outs[0][1], outs[1][0] + outs[1][1]):
            assert torch.equal(a, b), "peer-memory exchange differs from all_to_all_single"
        if rank == 0:
            print(f"[mgpu] NVLink peer-memory exchange == all_to_all_single (forward rows and backward gradients bit-exact, "
                  f"{outs[0][2][-1]} rows received on rank 0)")
    elif rank == 0:
        print("[mgpu] NVLink peer-memory exchange NOT available on this box: all_to_all_single path only")

    # ---- gather gradients ---------------------------------------------------------------------------
    names = ("_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity")
    grads = {}
    for n in names:
        g = getattr(tr.params, n).grad.contiguous()
        parts = [torch.empty_like(g) for _ in range(world)]  # equal shards: N divisible by world in this test
        dist.all_gather(parts, g)
        grads[n] = torch.cat(parts).cpu().numpy()
    losses = [torch.zeros((), device=dev) for _ in range(world)]
    dist.all_gather(losses, torch.tensor(loss, device=dev))

    if rank == 0:
        from oracle.oracle import Oracle
        # the other ranks wait in the next collective; torchrun's OMP_NUM_THREADS=1 default does not bind the oracle
        o = Oracle(np.float32, threads=max(1, min(32, (os.cpu_count() or 8) // 2)))
        T = tr.tile_y * tr.tile_x
        exp = {n: 0 for n in ("means3D", "scales", "rotations", "opacities", "shs")}
        tot_loss = 0.0
        for k, cam in enumerate(cams):
            pre = o.preprocess_forward(scene["means3D"], scene["scales"], scene["rotations"], scene["shs"], scene["opacities"], cam)
            fwd = o.render_forward(H, W, pre["means2D"], pre["conic_opacity"], pre["rgb"], pre["depths"], pre["radii"],
                                   np.ones(T, np.uint8), (0, 0, 0))
            err = np.abs(imgs[k] - fwd["image"])
            bad = (err > 1e-4 * np.abs(fwd["image"]) + 1e-5).mean()
            print(f"[mgpu] camera {k}: assembled image max_abs_err {err.max():.2e}, outside tol {bad:.1e}")
            assert bad <= 2e-4
            st = division.start_strategy([c["uid"] for c in cams], tr.history, world, 0)[0][k]
            dimg = np.zeros((3, H, W), np.float32)
            gtf = np.clip(gts[k].astype(np.float32) / np.float32(255), 0, 1)
            for li in range(len(st.gpu_ids)):
                y0, y1 = st.division_pos[li] * 16, min(st.division_pos[li + 1] * 16, H)
                l1, ss, g = o.loss(fwd["image"][:, y0:y1], gtf[:, y0:y1], H * W, 0.2)
                dimg[:, y0:y1] = g
                tot_loss += 0.8 * l1 + 0.2 * (1 - ss)
            rb = o.render_backward(H, W, pre["means2D"], pre["conic_opacity"], pre["rgb"], (0, 0, 0), fwd, dimg)
            pb = o.preprocess_backward(scene["means3D"], scene["scales"], scene["rotations"], scene["shs"],
                                       scene["opacities"], cam, pre["radii"], pre["clamped"], rb["means2D"],
                                       rb["conic_opacity"], rb["rgb"])
            for n in exp:
                exp[n] = exp[n] + pb[n]
        got_loss = float(sum(float(x) for x in losses))
        print(f"[mgpu] loss sum over ranks {got_loss:.6f} vs oracle strip-wise {tot_loss:.6f}")
        assert abs(got_loss - tot_loss) <= 1e-4 * abs(tot_loss)
        op, q = scene["opacities"], scene["rotations"]
        gq = exp["rotations"]
        ref = {"_xyz": exp["means3D"], "_features_dc": exp["shs"][:, :1], "_features_rest": exp["shs"][:, 1:],
               "_scaling": exp["scales"] * scene["scales"], "_opacity": exp["opacities"] * op * (1 - op),
               "_rotation": gq - q * (q * gq).sum(1, keepdims=True)}
        for n in names:
            a, b = grads[n].astype(np.float64), ref[n].astype(np.float64)
            rms = np.sqrt((b ** 2).mean())
            badf = (np.abs(a - b) > 1e-4 * np.abs(b) + 1e-4 * rms).mean()
            print(f"[mgpu] grad {n}: rms {rms:.3e} max_abs_err {np.abs(a - b).max():.3e} outside tol {badf:.1e}")
            assert badf <= 1e-3, n
        print("[mgpu] PASS world_size", world)

    # ---- 3. replicated-Gaussian gradient sync: fused sparse all-reduce == dense all-reduce -------------------
    from gs_b200 import grad_sync
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    P2 = 20011
    shapes = [(P2, 3), (P2, 1, 3), (P2, 15, 3), (P2, 3), (P2, 4), (P2, 1)]
    touched = torch.rand((P2,), device=dev, generator=g) < 0.07     # each rank touches ~7 % of the Gaussians
    params_a, params_b = [], []
    for shp in shapes:
        grad = torch.randn(shp, device=dev, generator=g)
        grad[~touched] = 0
        pa, pb = torch.zeros(shp, device=dev), torch.zeros(shp, device=dev)
        pa.grad, pb.grad = grad.clone(), grad.clone()
        params_a.append(pa); params_b.append(pb)
    n_touched = grad_sync.sync_gradients_fused_sparse(params_a)
    grad_sync.sync_gradients_densely(params_b)
    for pa, pb in zip(params_a, params_b):
        assert torch.allclose(pa.grad, pb.grad, rtol=1e-6, atol=1e-6), "fused sparse gradient sync != dense all-reduce"
    if rank == 0:
        print(f"[mgpu] fused sparse gradient all-reduce: {n_touched} of {P2} Gaussians touched, matches dense all-reduce")

    # ---- 4. border-pixel exchange (row L1): ONE camera split over all ranks; strip losses must add up to the
    #         FULL-image loss and the gradients must equal the single-GPU full-image gradients -----------------
    if tr.tile_y < 3 * world:  # 13 tile rows cannot be cut into `world` strips of >= 2 rows
        if rank == 0:
            print(f"[mgpu] border exchange: skipped at world_size {world} (image too short for {world} strips)")
        dist.barrier()
        dist.destroy_process_group()
        return
    tr2 = pipeline.Trainer(scene, cams[:1], [torch.from_numpy(gts[0]).pin_memory()], dev, rank, world, border_exchange=True)
    loss2 = tr2.step(resident=False)
    l2 = [torch.zeros((), device=dev) for _ in range(world)]
    dist.all_gather(l2, torch.tensor(loss2, device=dev))
    g2 = {}
    for n in names:
        g = getattr(tr2.params, n).grad.contiguous()
        parts = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(parts, g)
        g2[n] = torch.cat(parts).cpu().numpy()
    if rank == 0:
        ref2 = o.train_step(scene, cams[0], gts[0])
        # each strip adds lambda*(1 - its partial ssim): the constant lambda appears once per strip
        got = float(sum(float(x) for x in l2)) - 0.2 * (world - 1)
        print(f"[mgpu] border exchange: strip losses sum to {got:.6f}, full-image oracle loss {ref2['loss']:.6f}")
        assert abs(got - ref2["loss"]) <= 1e-4 * abs(ref2["loss"])
        e = ref2["grads"]
        gq2 = e["rotations"]
        refb = {"_xyz": e["means3D"], "_features_dc": e["shs"][:, :1], "_features_rest": e["shs"][:, 1:],
                "_scaling": e["scales"] * scene["scales"], "_opacity": e["opacities"] * op * (1 - op),
                "_rotation": gq2 - q * (q * gq2).sum(1, keepdims=True)}
        for n in names:
            a, b = g2[n].astype(np.float64), refb[n].astype(np.float64)
            rms = np.sqrt((b ** 2).mean())
            badf = (np.abs(a - b) > 1e-4 * np.abs(b) + 1e-4 * rms).mean()
            assert badf <= 1e-3, ("border exchange gradient", n, badf)
        print("[mgpu] border exchange: gradients equal the full-image (single-GPU) gradients: PASS")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
