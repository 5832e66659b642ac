#!/bin/bash
# One GPU-box visit for the event-split schedule: CE parity tests, then the same bench line with the fused kernel
# and with the event-split kernels at several slot counts.
# Usage (under gpurun, from the repo root): bash scripts/gpu_event_ab.sh TAG
set -u
TAG=${1:-ev}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest CE + host"
timeout 900 python -m pytest tests/test_gpu_ce.py tests/test_gpu_host.py -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_ce.txt
run() {  # name, env...
  local name=$1; shift
  echo "== bench $name"
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-multigroup 2>$OUT/$name.err | tee $OUT/$name.json | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   value %.4g  e2e %.4g  ms %.2f  launches %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']))"
  tail -3 $OUT/$name.err
}
run fused MMC_SCHEDULE=1
for s in ${SLOTS:-262144 1048576 2097152 4194304}; do
  run event_$s MMC_SCHEDULE=2 MMC_EVENT_SLOTS=$s
done
if [ -n "${CT:-}" ]; then
  echo "== continuous_temperature"
  for sch in 1 2; do
    MMC_SCHEDULE=$sch timeout 600 python bench.py --workload continuous_temperature --steps 3 --warmup 3 --no-cpu-baseline --no-multigroup 2>$OUT/ct_$sch.err | tee $OUT/ct_$sch.json | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ct schedule $sch value %.4g  ms %.2f' % (d['value'], d['ms_per_step']))"
  done
fi
if [ -n "${NCU:-}" ]; then
  echo "== ncu launch list (event schedule, 2^20 histories)"
  MMC_SCHEDULE=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_event.csv \
    python bench.py --steps 1 --warmup 1 --histories-per-gpu 1048576 --no-cpu-baseline --no-multigroup > $OUT/bench_under_ncu.log 2>&1
  echo "== ncu full: flight kernel and S(a,b) kernel of a steady-state pass"
  MMC_SCHEDULE=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:event_flight_kernel -s 6 -c 1 -o $OUT/prof_flight \
    python bench.py --steps 1 --warmup 0 --histories-per-gpu 8388608 --no-cpu-baseline --no-multigroup > $OUT/ncu_flight.log 2>&1
  MMC_SCHEDULE=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:event_tsl_kernel -s 6 -c 1 -o $OUT/prof_tsl \
    python bench.py --steps 1 --warmup 0 --histories-per-gpu 8388608 --no-cpu-baseline --no-multigroup > $OUT/ncu_tsl.log 2>&1
fi
ls -la $OUT
