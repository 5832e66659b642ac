#!/bin/bash
# On the GPU box: the CE k-eigenvalue bench with the base library and with the CEB2 variant (fused CE kernels at 2 CTAs
# per SM, scripts/build_variant.sh CEB2 -DMMC_CE_BLOCKS_PER_SM=2).  r02r: 2.285e7 against 2.048e7 hist/s: 3 CTAs stay.
set -u
OUT=gpurun_out/r02r; mkdir -p $OUT
cp minimc_b200/libminimc_b200.so /tmp/base.so
for v in base CEB2; do
  [ $v = base ] || cp minimc_b200/csrc/build/variants/$v.so minimc_b200/libminimc_b200.so
  python bench.py --workload keigenvalue_ce --steps 6 --warmup 2 2>/dev/null > $OUT/kce_$v.json
  python -c "import json;j=json.loads(open('$OUT/kce_$v.json').read().strip().splitlines()[-1]);print('$v', '%.4g'%j['value'], j['ms_per_step'])"
done
cp /tmp/base.so minimc_b200/libminimc_b200.so
