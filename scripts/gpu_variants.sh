#!/bin/bash
# On the GPU box: for every minimc_b200/csrc/build/variants/*.so, swap it in, run the CE parity tests and a short bench.
# Usage: bash scripts/gpu_variants.sh TAG [bench args...]
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
cp minimc_b200/libminimc_b200.so /tmp/libminimc_b200.base.so
for v in base minimc_b200/csrc/build/variants/*.so; do
  if [ "$v" = base ]; then name=base; cp /tmp/libminimc_b200.base.so minimc_b200/libminimc_b200.so
  else [ -f "$v" ] || continue; name=$(basename $v .so); cp $v minimc_b200/libminimc_b200.so; fi
  echo "== variant $name"
  timeout 600 python -m pytest tests/test_gpu_ce.py tests/test_gpu_full_shape.py -m gpu -x -q 2>&1 | tail -2 | tee $OUT/$name.pytest.txt
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-multigroup --no-extras "$@" 2>$OUT/$name.err | tee $OUT/$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   value %.4g  e2e %.4g  ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
cp /tmp/libminimc_b200.base.so minimc_b200/libminimc_b200.so
