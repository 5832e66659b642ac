#!/bin/bash
# One GPU-box visit for the dense reconstruction tables: full GPU parity with the tables, CE parity without them
# (MMC_TSL_DENSE_MB=0: the on-the-fly sums), then the bench line for each and for every variant .so.
# Usage (under gpurun): bash scripts/gpu_dense_ab.sh TAG
set -u
TAG=${1:-dense}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu (dense tables)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_dense.txt
echo "== pytest CE (MMC_TSL_DENSE_MB=0)"; MMC_TSL_DENSE_MB=0 timeout 900 python -m pytest tests/test_gpu_ce.py tests/test_gpu_host.py -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_nodense.txt
run() {
  local name=$1; shift
  echo "== bench $name"
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-multigroup ${BENCH_ARGS:-} 2>$OUT/$name.err | tee $OUT/$name.json | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); ks=d['roofline']['kernel_split']; print('   value %.4g  e2e %.4g  ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']), {k: round(v, 2) for k, v in ks.items() if k.endswith('_ms')})"
  tail -3 $OUT/$name.err
}
run dense_base MMC_X=1
run nodense MMC_TSL_DENSE_MB=0
cp minimc_b200/libminimc_b200.so /tmp/base.so
for v in minimc_b200/csrc/build/variants/*.so; do
  [ -f "$v" ] || continue
  cp $v minimc_b200/libminimc_b200.so
  run $(basename $v .so) MMC_X=1
done
cp /tmp/base.so minimc_b200/libminimc_b200.so
for wl in continuous_temperature multi_zone broomstick; do
  echo "== $wl"
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-multigroup 2>$OUT/$wl.err | tee $OUT/$wl.json | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   value %.4g  e2e %.4g  ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
