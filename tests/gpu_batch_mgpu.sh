#!/bin/bash
# One gpurun --gpus N call: multi-GPU parity log, the headline bench line at N ranks (with its untimed strong-scaling
# extras and parity verdict), and the BASELINE.json configuration that belongs to N (c3 at 4, c4 + c5 at 8).
# usage: bash tests/gpu_batch_mgpu.sh <N> <tag>
N=${1:-2}; TAG=${2:-r2x}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout ${PARITY_TIMEOUT:-300} $TR --master-port 29511 tests/mgpu_parity.py > $OUT/${TAG}_mgpu_parity_${N}gpu.log 2>&1; echo "parity rc=$?"
tail -4 $OUT/${TAG}_mgpu_parity_${N}gpu.log
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 4), d.get("phase_ms_serialised"),
          {k: round(v["ms_per_step"], 3) for k, v in (d.get("extra") or {}).items()}, (d["config"].get("parity") or {}).get("ok"))
except Exception as e:
    print("no line:", e)
PY
}
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_${N}gpu.json \
    2> $OUT/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"; show $OUT/${TAG}_bench_${N}gpu.json; tail -2 $OUT/${TAG}_bench_${N}gpu.err
if [ -n "$GAPS" ]; then
  timeout 300 $TR --master-port 29515 tests/gap_profile.py --views $N --cprofile > $OUT/${TAG}_gaps_${N}gpu.log 2>&1; echo "gaps rc=$?"
  head -30 $OUT/${TAG}_gaps_${N}gpu.log | cut -c1-200
fi
if [ -n "$STRONG" ]; then
  timeout 300 $TR --master-port 29516 tests/gap_profile.py --views 1 --cprofile > $OUT/${TAG}_gaps_strong_${N}gpu.log 2>&1; echo "strong gaps rc=$?"
  grep -A28 "host hot spots" $OUT/${TAG}_gaps_strong_${N}gpu.log | cut -c1-160
  grep -m3 "host enqueue\|last full step" $OUT/${TAG}_gaps_strong_${N}gpu.log
fi
if [ "$N" = 4 ] && [ -z "$SKIP_CFG" ]; then
  timeout 600 $TR --master-port 29513 bench.py --gpus 4 --workload c3 --steps 20 --warmup 5 --no-cpu-baseline --time-optimizer \
      > $OUT/${TAG}_bench_c3_4gpu.json 2> $OUT/${TAG}_bench_c3_4gpu.err; echo "c3 rc=$?"; show $OUT/${TAG}_bench_c3_4gpu.json
  tail -2 $OUT/${TAG}_bench_c3_4gpu.err
fi
if [ "$N" = 8 ] && [ -z "$SKIP_CFG" ]; then
  timeout 900 $TR --master-port 29513 bench.py --gpus 8 --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extra \
      --time-optimizer > $OUT/${TAG}_bench_c4_8gpu.json 2> $OUT/${TAG}_bench_c4_8gpu.err; echo "c4 rc=$?"
  show $OUT/${TAG}_bench_c4_8gpu.json; tail -2 $OUT/${TAG}_bench_c4_8gpu.err
  if [ -z "$SKIP_C5" ]; then
    timeout 900 $TR --master-port 29514 tests/c5_stress.py --iters ${C5_ITERS:-1000} > $OUT/${TAG}_c5_8gpu.log 2>&1; echo "c5 rc=$?"
    tail -3 $OUT/${TAG}_c5_8gpu.log
  fi
fi
