"""One-call A/B of the experimental kernels (include/grendel_gs_b200.h, gs_debug_set) on a B200:

    python tests/ab_experimental.py            # parity of every variant, then bench.py per variant

1. runs the gated parity tests (GS_B200_EXPERIMENTAL=1) -- a variant that fails parity is not timed;
2. runs `bench.py --no-cpu-baseline` once per variant (GS_B200_DEBUG_FLAGS) and prints ms/step plus the per-stage
   device times, so one gpurun call answers "which backward kernel should be the default".
Diagnostics only: numbers from here are not bench values (the shipped configuration is flags = 0).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = [("default", 0), ("bwd_wht_64", 2), ("bwd_wht_128", 4), ("bwd_auto", 8)]


def main():
    env = dict(os.environ, GS_B200_EXPERIMENTAL="1")
    ok = {}
    # 0. everything else that was written without device access: fused Adam, fused densification step
    for f in ("test_zz_fused_adam_gpu.py", "test_zz_densify_gpu.py", "test_zz_knn_gpu.py"):
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", f), "-q", "-m", "gpu"], env=env,
                           capture_output=True, text=True, cwd=ROOT)
        print(f"[ab] {f}: {'PASS' if r.returncode == 0 else 'FAIL'}  {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ''}",
              flush=True)
        if r.returncode != 0:
            print(r.stdout[-3000:], r.stderr[-2000:], flush=True)
    for name, flag in VARIANTS[1:]:
        key = {2: "DEBUG_BWD_WHT_64", 4: "DEBUG_BWD_WHT_128", 8: "DEBUG_BWD_AUTO"}[flag]
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x",
                            "-m", "gpu", "-k", f"experimental and {key}"], env=env, capture_output=True, text=True, cwd=ROOT)
        ok[name] = r.returncode == 0
        print(f"[ab] parity {name}: {'PASS' if ok[name] else 'FAIL'}", flush=True)
        if not ok[name]:
            print(r.stdout[-3000:], r.stderr[-2000:], flush=True)
    rows = []
    for name, flag in VARIANTS:
        if name != "default" and not ok.get(name, False):
            continue
        e = dict(os.environ, GS_B200_DEBUG_FLAGS=str(flag))
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--steps", "30", "--warmup", "5"]
        if name == "default":
            cmd.append("--time-optimizer")
        r = subprocess.run(cmd, env=e, capture_output=True, text=True, cwd=ROOT)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not line:
            print(f"[ab] bench {name}: FAILED\n{r.stderr[-2000:]}", flush=True)
            continue
        d = json.loads(line[-1])
        st = d["roofline"]["stage_ms_per_launch"]
        rows.append((name, d["ms_per_step"], d["e2e"]["ms_per_step"], st.get("b10 render"), st.get("70 render")))
        print(f"[ab] {name}: {d['ms_per_step']:.3f} ms/step (e2e {d['e2e']['ms_per_step']:.3f}); b10 render {st.get('b10 render')} ms; "
              f"loss_check {d['config']['loss_check']:.6f}", flush=True)
        if "optimizer" in d:
            o = d["optimizer"]
            print(f"[ab] fused Adam: {o['ms_per_step']:.3f} ms/step, {o['achieved_gbs']:.0f} GB/s "
                  f"({100 * o['frac_of_hbm_peak']:.0f} % of measured copy bandwidth)", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ab_experimental.json"), "w") as f:
        json.dump([dict(variant=n, ms_per_step=a, e2e_ms=b, bwd_ms=c, fwd_ms=dd) for n, a, b, c, dd in rows], f, indent=1)


if __name__ == "__main__":
    main()
