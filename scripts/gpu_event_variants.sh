#!/bin/bash
# On the GPU box: base library + every minimc_b200/csrc/build/variants/*.so: CE parity tests, then the bench line with
# the event-split schedule.  Usage: bash scripts/gpu_event_variants.sh TAG [bench args...]
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
cp minimc_b200/libminimc_b200.so /tmp/libminimc_b200.base.so
for v in base minimc_b200/csrc/build/variants/*.so; do
  if [ "$v" = base ]; then name=base; cp /tmp/libminimc_b200.base.so minimc_b200/libminimc_b200.so
  else [ -f "$v" ] || continue; name=$(basename $v .so); cp $v minimc_b200/libminimc_b200.so; fi
  echo "== variant $name"
  timeout 600 python -m pytest tests/test_gpu_ce.py -m gpu -x -q 2>&1 | tail -2 | tee $OUT/$name.pytest.txt
  MMC_SCHEDULE=2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-multigroup "$@" 2>$OUT/$name.err | tee $OUT/$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   value %.4g  e2e %.4g  ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
  tail -2 $OUT/$name.err
done
cp /tmp/libminimc_b200.base.so minimc_b200/libminimc_b200.so
if [ -n "${NCU:-}" ]; then
  echo "== ncu launch list (event schedule, 2^21 histories)"
  MMC_SCHEDULE=2 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches_event.csv \
    python bench.py --steps 1 --warmup 0 --histories-per-gpu 2097152 --no-cpu-baseline --no-multigroup > $OUT/bench_under_ncu.log 2>&1
  MMC_SCHEDULE=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:event_flight_kernel -s 6 -c 1 -o $OUT/prof_flight \
    python bench.py --steps 1 --warmup 0 --histories-per-gpu 8388608 --no-cpu-baseline --no-multigroup > $OUT/ncu_flight.log 2>&1
  MMC_SCHEDULE=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:event_tsl_kernel -s 6 -c 1 -o $OUT/prof_tsl \
    python bench.py --steps 1 --warmup 0 --histories-per-gpu 8388608 --no-cpu-baseline --no-multigroup > $OUT/ncu_tsl.log 2>&1
fi
