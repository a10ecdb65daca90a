"""One-call A/B of the two backward blend kernels (include/grendel_gs_b200.h, gs_debug_set) on a B200:

    python tests/ab_experimental.py

runs `bench.py --no-cpu-baseline` once per variant (GS_B200_DEBUG_FLAGS) and prints ms/step plus the per-stage device
times.  Diagnostics only: numbers from here are not bench values (the shipped configuration is flags = 0).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = [("segment-parallel (default)", 0), ("tile-parallel (round 1)", 2)]


def main():
    rows = []
    for name, flag in VARIANTS:
        e = dict(os.environ, GS_B200_DEBUG_FLAGS=str(flag))
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--steps", "30", "--warmup", "5"]
        r = subprocess.run(cmd, env=e, capture_output=True, text=True, cwd=ROOT)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not line:
            print(f"[ab] bench {name}: FAILED\n{r.stdout[-2000:]}\n{r.stderr[-3000:]}", flush=True)
            continue
        d = json.loads(line[-1])
        st = d["roofline"]["stage_ms_per_launch"]
        rows.append(dict(variant=name, ms_per_step=d["ms_per_step"], e2e_ms=d["e2e"]["ms_per_step"], stages=st))
        print(f"[ab] {name}: {d['ms_per_step']:.3f} ms/step (e2e {d['e2e']['ms_per_step']:.3f}); "
              f"loss_check {d['config']['loss_check']:.6f}\n     stages: "
              + ", ".join(f"{k} {v:.3f}" for k, v in st.items()), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ab_experimental.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
