#!/bin/bash
# Sweep of the hand-over threshold (live histories at or below which the drain goes to the fused kernel).
# Usage (under gpurun): bash scripts/gpu_handover_sweep.sh TAG [thresholds...]
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for h in "$@"; do
  for wl in single_zone multi_zone; do
  MMC_EVENT_HANDOVER=$h timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-multigroup 2>$OUT/h$h.$wl.err | tee $OUT/h$h.$wl.json | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); ks=d['roofline']['kernel_split']; print('handover $h $wl value %.4g  ms %.2f launches %s' % (d['value'], d['ms_per_step'], d['gpu_launches']), {k: round(v, 2) for k, v in ks.items() if k.endswith('_ms')})"
  done
done
