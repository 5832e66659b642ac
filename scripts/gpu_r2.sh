#!/bin/bash
# One GPU-box visit of round 2.  Usage (under gpurun, from the repo root): bash scripts/gpu_r2.sh TAG [stage ...]
# stages: test smoke bench bench_all ref launches ncu_tsl ncu_flight small strong
set -u
TAG=${1:-r02}; shift
STAGES=${*:-test smoke bench}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
for st in $STAGES; do case $st in
test) echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt;;
testce) echo "== pytest CE"; timeout 1500 python -m pytest tests/test_gpu_ce.py tests/test_gpu_full_shape.py -x -q 2>&1 | tail -25 | tee $OUT/pytest_ce.txt;;
smoke) echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt;;
ref) echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>$OUT/bench_ref.err | tee $OUT/bench_ref.json;;
bench) echo "== bench"; timeout 1500 python bench.py --gpus 1 --steps 10 --warmup 3 ${BENCH_ARGS:-} 2>$OUT/bench.err | tee $OUT/bench.json;;
benchq) echo "== bench (quick)"; timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-multigroup --no-cpu-baseline --no-extras ${BENCH_ARGS:-} 2>$OUT/benchq.err | tee $OUT/benchq.json;;
bench_all) for wl in continuous_temperature multi_zone broomstick; do
  echo "== bench $wl"; timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 --no-multigroup --no-cpu-baseline 2>$OUT/bench_$wl.err | tee $OUT/bench_$wl.json; done;;
small) for n in 1000000 4194304 12500000; do
  echo "== bench single_zone at $n histories"; timeout 900 python bench.py --steps 10 --warmup 3 --histories-per-gpu $n --no-multigroup --no-cpu-baseline 2>$OUT/bench_n$n.err | tee $OUT/bench_n$n.json; done;;
keig) for wl in keigenvalue_mg keigenvalue_ce; do
  echo "== bench $wl"; timeout 900 python bench.py --workload $wl --steps 8 --warmup 3 2>$OUT/bench_$wl.err | tee $OUT/bench_$wl.json; done;;
strong) echo "== bench strong scaling 1e8"; timeout 900 python bench.py --scaling strong --total-histories 100000000 --steps 5 --warmup 3 --no-multigroup --no-cpu-baseline 2>$OUT/bench_strong.err | tee $OUT/bench_strong.json;;
launches) echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --histories-per-gpu 4194304 --no-cpu-baseline --no-multigroup > $OUT/bench_under_ncu.log 2>&1;;
ncu_tsl) echo "== ncu full: S(a,b) kernel of a steady-state pass"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:event_tsl -s ${NCU_SKIP:-6} -c 1 -o $OUT/prof_tsl \
    python bench.py --steps 1 --warmup 0 --histories-per-gpu ${NCU_HIST:-16777216} --no-cpu-baseline --no-multigroup --no-extras > $OUT/ncu_tsl.log 2>&1;;
ncu_flight) echo "== ncu full: flight kernel of a steady-state pass"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:event_flight_kernel -s ${NCU_SKIP:-6} -c 1 -o $OUT/prof_flight \
    python bench.py --steps 1 --warmup 0 --histories-per-gpu ${NCU_HIST:-16777216} --no-cpu-baseline --no-multigroup --no-extras > $OUT/ncu_flight.log 2>&1;;
slots) for n in 1048576 4194304; do
  echo "== bench single_zone with $n slots"; MMC_EVENT_SLOTS=$n timeout 900 python bench.py --steps 5 --warmup 3 --no-multigroup --no-cpu-baseline --no-extras 2>$OUT/bench_slots$n.err | tee $OUT/bench_slots$n.json; done;;
*) echo "unknown stage $st";;
esac; done
ls -la $OUT
