"""BASELINE.json configs[4] ("densification stress: 1M -> 20M Gaussians over 1000 iters, batch-size 4, 8 x B200"), run
under torch.distributed.run, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tests/c5_stress.py [--iters 1000] [--n0 1000000] [--n1 20000000] [--views 4] [--width 1920 --height 1080]

The reference's iteration (train_internal.py:139-340): step (preprocess -> all-to-all -> render -> loss -> backward),
densification statistics (densification.py:14-24), optimizer step with gradients / bsz (fused Adam, one launch), and every
100 iterations: densify_and_prune (fused select / scan / gather kernels), host-side growth to the scheduled total
("append freshly sampled Gaussians every 100 it; host-side"), redistribution when the shards are uneven
(one fused all-to-all).  Prints ONE JSON line on rank 0 with the per-phase times; diagnostics, not a bench value.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "grendel-gs_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from gs_b200 import densify, pipeline, redistribute, synthetic as syn  # noqa: E402
from gs_b200.optim import FusedAdam                                     # noqa: E402


def raw_of(scene):
    """Activated synthetic Gaussians -> the raw parameterisation of the six groups."""
    t = lambda a: torch.as_tensor(a, dtype=torch.float32)
    op = t(scene["opacities"]).clamp(1e-6, 1 - 1e-6)
    shs = t(scene["shs"])
    return {"xyz": t(scene["means3D"]), "f_dc": shs[:, :1].contiguous(), "f_rest": shs[:, 1:].contiguous(),
            "opacity": torch.log(op / (1 - op)), "scaling": torch.log(t(scene["scales"])), "rotation": t(scene["rotations"])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--n0", type=int, default=1_000_000)
    ap.add_argument("--n1", type=int, default=20_000_000)
    ap.add_argument("--views", type=int, default=4)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--interval", type=int, default=100)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    W, H, B = a.width, a.height, a.views
    lo, hi = a.n0 * rank // world, a.n0 * (rank + 1) // world
    scene = syn.make_scene_shard(a.n0, lo, hi, W, H, seed=3)
    cams = syn.make_batch_cameras(W, H, B)
    gts = [torch.from_numpy(syn.make_gt_image(W, H, seed=1 + k)).pin_memory() for k in range(B)]
    tr = pipeline.Trainer(scene, cams, gts, dev, rank, world, shard=(lo, hi, a.n0), peer_cap_rows=int(1.3 * a.n1) + 65536)
    opt = FusedAdam(tr.optimizer_groups(), lr=0.0, eps=1e-15)
    P = tr.n_local
    accum, denom = torch.zeros((P,), device=dev), torch.zeros((P,), device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t_step, t_opt, t_dens, t_red, sizes, moved = [], [], [], [], [], 0
    n_appended = 0
    wall0 = time.perf_counter()
    for it in range(1, a.iters + 1):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        tr.step(resident=True)
        e1.record()
        with torch.no_grad():     # add_densification_stats (gaussian_model.py:1046-1071) over the B cameras
            g = tr.means2D.grad if not isinstance(tr.means2D, list) else torch.stack([m.grad for m in tr.means2D])
            vis = tr._radii_local > 0
            accum += (g.norm(dim=-1) * vis).sum(dim=0)
            denom += vis.sum(dim=0)
        opt.step(grad_scale=1.0 / B)
        e2.record()
        if it % 10 == 0 or it == 1:
            torch.cuda.synchronize()
            t_step.append(e0.elapsed_time(e1)); t_opt.append(e1.elapsed_time(e2))
        if it % a.interval == 0:
            torch.cuda.synchronize()
            d0 = time.perf_counter()
            res = densify.densify_and_prune(opt, accum.reshape(-1, 1), denom.reshape(-1, 1), 0.0002, 0.005, 10.0, 0.01, None)
            torch.cuda.synchronize()
            d1 = time.perf_counter()
            # host-side growth to the scheduled total: fresh Gaussians, this rank's share
            n_now = torch.tensor([res["xyz"].shape[0]], dtype=torch.int64, device=dev)
            dist.all_reduce(n_now)
            target = a.n0 + (a.n1 - a.n0) * it // a.iters
            grow = max(0, target - int(n_now.item()))
            mine = grow * (rank + 1) // world - grow * rank // world
            new = res
            if mine > 0:
                fresh = syn.make_scene_shard(a.n1 * 4, n_appended + lo % 7, n_appended + lo % 7 + mine, W, H, seed=1000 + rank)
                new = densify.append_gaussians(opt, raw_of(fresh))
                n_appended += mine
            tr.adopt_parameters({k: new[k] for k in densify.NAMES})
            torch.cuda.synchronize()
            d2 = time.perf_counter()
            need, counts = redistribute.need_redistribute(tr.n_local, threshold=1.05, first_after_densify=(it == a.interval))
            if need:
                r = redistribute.redistribute(opt)
                tr.adopt_parameters({k: r[k] for k in densify.NAMES})
                moved += 1
            torch.cuda.synchronize()
            d3 = time.perf_counter()
            t_dens.append((d1 - d0) * 1e3); t_red.append((d3 - d2) * 1e3)
            P = tr.n_local
            accum, denom = torch.zeros((P,), device=dev), torch.zeros((P,), device=dev)
            tot = torch.tensor([P], dtype=torch.int64, device=dev)
            dist.all_reduce(tot)
            sizes.append(int(tot.item()))
            if rank == 0:
                print(f"[c5] it {it}: {sizes[-1]} Gaussians (densify kept/clones/split {res['counts'][:4]}), step "
                      f"{np.median(t_step[-10:]):.2f} ms, adam {np.median(t_opt[-10:]):.2f} ms, densify {t_dens[-1]:.1f} ms, "
                      f"redistribute {'%.1f ms' % t_red[-1] if need else 'no'}; shard sizes {counts}", flush=True)
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    mem = torch.cuda.max_memory_allocated() / 2**30
    memt = torch.tensor([mem], device=dev)
    dist.all_reduce(memt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": f"c5: {a.n0} -> {sizes[-1] if sizes else a.n0} Gaussians over {a.iters} iterations, "
                                      f"bsz {B}, {world} GPUs, {W}x{H}",
                          "wall_s": wall, "iterations_per_s": a.iters / wall, "step_ms_first": t_step[0],
                          "step_ms_last": float(np.median(t_step[-10:])), "adam_ms_last": float(np.median(t_opt[-10:])),
                          "densify_ms": t_dens, "redistribute_ms": t_red, "redistributions": moved, "sizes": sizes,
                          "max_memory_GiB_per_gpu": float(memt.item())}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
