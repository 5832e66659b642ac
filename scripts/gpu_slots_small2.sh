#!/bin/bash
# Finer slot-count sweep around 10^6 histories (the deck as shipped), after scripts/gpu_slots_small.sh:
# bash scripts/gpu_slots_small2.sh TAG   (r02x: 2^19 slots best at 10^6 and 2 * 10^6 histories)
OUT=gpurun_out/$1; mkdir -p $OUT
run() { MMC_EVENT_SLOTS=$2 python bench.py --steps 10 --warmup 3 --histories-per-gpu $1 --no-multigroup --no-cpu-baseline --no-extras 2>/dev/null > $OUT/n$1_s$2.json
  python -c "import json;j=json.loads(open('$OUT/n$1_s$2.json').read().strip().splitlines()[-1]);k=j['roofline']['kernel_split'];print('n=$1 slots=$2', '%.4g'%j['value'], 'ms %.2f'%j['ms_per_step'], 'launches', k['launches'])"; }
for s in 131072 262144 393216 524288 786432; do run 1000000 $s; done
for s in 262144 524288 1048576; do run 2000000 $s; done
for s in 131072 262144 524288; do run 500000 $s; done
