"""One-call A/B of kernel variants on a B200 (diagnostics only: numbers from here are not bench values).

    python tests/ab_experimental.py                       # segment-parallel vs tile-parallel backward (gs_debug_set)
    python tests/ab_experimental.py name=variant[:flags] ...   # tuning builds: lib/libgrendel_gs_b200.<variant>.so
                                                               # (gs_b200.build.build(defs=[...], variant=...)), "default" = shipped

runs `bench.py --no-cpu-baseline` once per variant (GS_B200_LIB / GS_B200_DEBUG_FLAGS) and prints ms/step plus the
per-stage device times.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "grendel-gs_b200", "lib")


def main():
    variants = [("segment-parallel (default)", None, 0), ("tile-parallel (round 1)", None, 2)]
    if len(sys.argv) > 1:
        variants = []
        for a in sys.argv[1:]:
            name, _, rest = a.partition("=")
            var, _, flags = (rest or name).partition(":")
            lib = None if var == "default" else os.path.join(LIBDIR, f"libgrendel_gs_b200.{var}.so")
            variants.append((name, lib, int(flags or 0)))
    rows = []
    for name, lib, flag in variants:
        e = dict(os.environ, GS_B200_DEBUG_FLAGS=str(flag))
        if lib:
            e["GS_B200_LIB"] = lib
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
              f"loss_check {d['config']['loss_check']:.6f}; 70 render {st.get('70 render')}; b10 render {st.get('b10 render')}", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ab_experimental.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
