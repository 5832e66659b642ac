#!/bin/bash
# Slot-count sweep at the batch sizes below the large-batch threshold: bash scripts/gpu_slots_small.sh TAG
OUT=gpurun_out/$1; mkdir -p $OUT
for n in 1000000 4194304 12500000; do for s in 524288 1048576 2097152; do
  MMC_EVENT_SLOTS=$s python bench.py --steps 8 --warmup 3 --histories-per-gpu $n --no-multigroup --no-cpu-baseline --no-extras 2>/dev/null > $OUT/n${n}_s$s.json
  python -c "import json;j=json.loads(open('$OUT/n${n}_s$s.json').read().strip().splitlines()[-1]);k=j['roofline']['kernel_split'];print('n=$n slots=$s', '%.4g'%j['value'], 'ms %.2f'%j['ms_per_step'], 'launches', k['launches'])"
done; done
