#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the reference arm, the ncu launch list and full captures.
# Usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
if [ -z "${SKIP_REF:-}" ]; then
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>$OUT/bench_ref.err | tee $OUT/bench_ref.json
fi
echo "== bench"; timeout 1500 python bench.py --gpus 1 --steps 10 --warmup 3 ${BENCH_ARGS:-} 2>$OUT/bench.err | tee $OUT/bench.json
echo "== bench continuous_temperature"; timeout 900 python bench.py --workload continuous_temperature --steps 5 --warmup 3 --no-multigroup --no-cpu-baseline 2>$OUT/bench_ct.err | tee $OUT/bench_ct.json
for wl in multi_zone broomstick; do
echo "== bench $wl"; timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 --no-multigroup --no-cpu-baseline 2>$OUT/bench_$wl.err | tee $OUT/bench_$wl.json
done
echo "== bench fused schedule (for comparison)"; MMC_SCHEDULE=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-multigroup --no-cpu-baseline 2>$OUT/bench_fused.err | tee $OUT/bench_fused.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 1 --warmup 1 --histories-per-gpu 4194304 --no-cpu-baseline --no-multigroup > $OUT/bench_under_ncu.log 2>&1
echo "== ncu full (S(a,b) kernel and flight kernel of a steady-state pass)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:event_tsl_kernel -s 6 -c 1 -o $OUT/prof_tsl \
  python bench.py --steps 1 --warmup 0 --histories-per-gpu 8388608 --no-cpu-baseline --no-multigroup > $OUT/ncu_tsl.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:event_flight_kernel -s 6 -c 1 -o $OUT/prof_flight \
  python bench.py --steps 1 --warmup 0 --histories-per-gpu 8388608 --no-cpu-baseline --no-multigroup > $OUT/ncu_flight.log 2>&1
if [ -z "${SKIP_MG:-}" ]; then
echo "== ncu full (MG kernel)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fixed_source_kernel -s 1 -c 1 -o $OUT/prof_mg \
  python bench.py --workload multigroup_critical --steps 2 --warmup 1 --histories-per-gpu 134217728 --no-cpu-baseline > $OUT/bench_under_ncu_full_mg.log 2>&1
fi
ls -la $OUT
