"""Diagnostics: where the GPU idles inside a training step.  Runs Trainer.step under torch.profiler (CUPTI kernel
timeline), takes the last steps and prints every gap between consecutive GPU activities longer than 4 us with the
activities on both sides, plus busy / idle totals per step.  No nsys in the image; this is the substitute.

    python tests/gap_profile.py [--views 1] [--steps 6]            (one GPU; writes gpurun_out/gap_profile_*.txt)
    torchrun --nproc-per-node N tests/gap_profile.py --views N     (N ranks; rank 0's timeline is reported)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "grendel-gs_b200")]
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from gs_b200 import pipeline, synthetic as syn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--e2e", action="store_true")
    ap.add_argument("--cprofile", action="store_true", help="also print the host-side hot spots of 30 steps (cProfile)")
    a = ap.parse_args()
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, H, N = 1920, 1080, 2_000_000
    scene = syn.make_scene(N, W, H)
    cams = syn.make_batch_cameras(W, H, a.views)
    gts = [torch.from_numpy(syn.make_gt_image(W, H, seed=1 + k)).pin_memory() for k in range(a.views)]
    tr = pipeline.Trainer(scene, cams, gts, dev, rank, world)
    for _ in range(5):
        tr.step(resident=not a.e2e)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(20):
        tr.step(resident=not a.e2e)
    t_host = (time.perf_counter() - t0) / 20     # host time to ENQUEUE a step (no sync inside the loop except R)
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 20
    if a.cprofile and rank == 0:
        import cProfile, io, pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(30):
            tr.step(resident=not a.e2e)
        pr.disable()
        torch.cuda.synchronize()
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(35)
        print("host hot spots over 30 steps (cProfile, tottime):\n" + buf.getvalue()[:6000])
    elif a.cprofile:
        for _ in range(30):
            tr.step(resident=not a.e2e)
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            tr.step(resident=not a.e2e)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    lines = [f"host enqueue time per step {t_host * 1e3:.3f} ms; wall per step {t_all * 1e3:.3f} ms ({a.views} view(s))"]
    # split into steps at the preprocess kernel
    starts = [i for i, e in enumerate(evs) if "k_preprocess_fwd" in e.name]
    if len(starts) >= 3:
        lo, hi = starts[-2], starts[-1]
        step = evs[lo:hi]
        t_begin, t_end = step[0].time_range.start, evs[hi].time_range.start
        busy = sum(e.time_range.end - e.time_range.start for e in step)
        lines.append(f"last full step: {len(step)} GPU activities, span {(t_end - t_begin):.1f} us, busy {busy:.1f} us, "
                     f"idle {(t_end - t_begin - busy):.1f} us")
        prev = step[0]
        for e in step[1:] + [evs[hi]]:
            gap = e.time_range.start - prev.time_range.end
            if gap > 4.0:
                lines.append(f"  gap {gap:7.1f} us after {prev.name[:60]!r} ({prev.time_range.end - prev.time_range.start:.1f} us) "
                             f"before {e.name[:60]!r}")
            prev = e
        lines.append("activities of the step (us):")
        for e in step:
            lines.append(f"  {e.time_range.start - t_begin:9.1f} +{e.time_range.end - e.time_range.start:8.1f}  {e.name[:90]}")
    out = "\n".join(lines)
    if rank == 0:
        print(out)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"gap_profile_{world}gpu_{a.views}v{'_e2e' if a.e2e else ''}.txt"), "w") as f:
            f.write(out + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
