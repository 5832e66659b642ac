#!/bin/bash
# A/B of the persisting-L2 window on one box: bash scripts/gpu_l2_ab.sh TAG
OUT=gpurun_out/$1; mkdir -p $OUT
for rep in 1 2; do for v in 0 1; do
  MMC_L2_PERSIST=$v python bench.py --steps 5 --warmup 3 --no-multigroup --no-cpu-baseline --no-extras 2>/dev/null > $OUT/l2_${v}_$rep.json
  python -c "import json;j=json.loads(open('$OUT/l2_${v}_$rep.json').read().strip().splitlines()[-1]);k=j['roofline']['kernel_split'];print('persist=$v rep=$rep', '%.4g'%j['value'], '%.1f %.1f'%(k['event_flight_kernel_ms'],k['event_tsl_kernel_ms']))"
done; done
for v in 0 1; do MMC_L2_PERSIST=$v python bench.py --workload multi_zone --steps 4 --warmup 2 --no-multigroup --no-cpu-baseline --no-extras 2>/dev/null > $OUT/l2mz_$v.json
  python -c "import json;j=json.loads(open('$OUT/l2mz_$v.json').read().strip().splitlines()[-1]);print('multi_zone persist=$v', '%.4g'%j['value'])"; done
