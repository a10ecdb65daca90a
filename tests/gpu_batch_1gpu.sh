#!/bin/bash
# One gpurun call on ONE B200: device suite, bench A/B of library variants, ncu captures of the two blend kernels,
# single-rank sanity runs of the multi-GPU harnesses.  Everything lands in gpurun_out/ (tag = $1).
TAG=${1:-r2x}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/${TAG}_pytest_gpu.log; grep "c2:" $OUT/${TAG}_pytest_gpu.log
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["ms_per_step"], 4), round(d["e2e"]["ms_per_step"], 4), d["roofline"]["stage_ms_per_launch"])
PY
}
for v in default ${VARIANTS}; do
  L=""; [ $v != default ] && L=$PWD/grendel-gs_b200/lib/libgrendel_gs_b200.$v.so
  EXTRA="--no-cpu-baseline"; [ $v = default ] && EXTRA="--time-optimizer"
  GS_B200_LIB=$L timeout 300 python bench.py --steps 30 --warmup 5 $EXTRA > $OUT/${TAG}_bench_$v.json 2> $OUT/${TAG}_bench_$v.err
  echo "bench $v rc=$?"; show $OUT/${TAG}_bench_$v.json
done
if [ -n "$NCU" ]; then
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_blend -c 2 -f -o $OUT/${TAG}_blend \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_blend.log 2>&1; echo "ncu rc=$?"
fi
if [ -n "$LAUNCHES" ]; then   # the ncu launch list of B200_PROFILING.md: per-launch GPU time of every kernel of the bench command
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_ncu_gputime.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1; echo "launch list rc=$?"
fi
if [ -n "$GAPS" ]; then
  timeout 300 python tests/gap_profile.py --cprofile > $OUT/${TAG}_gaps.log 2>&1; echo "gaps rc=$?"; head -12 $OUT/${TAG}_gaps.log
fi
if [ -n "$SANITY" ]; then
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1"
  timeout 300 $TR --master-port 29541 tests/c5_stress.py --iters 300 --n0 200000 --n1 1500000 --views 2 --interval 100 \
      > $OUT/${TAG}_c5_1gpu.log 2>&1; echo "c5 rc=$?"; tail -4 $OUT/${TAG}_c5_1gpu.log
  timeout 300 python bench.py --workload c3 --views 1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_c3_1gpu.json \
      2> $OUT/${TAG}_bench_c3_1gpu.err; echo "c3 rc=$?"; show $OUT/${TAG}_bench_c3_1gpu.json
fi
