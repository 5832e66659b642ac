#!/bin/bash
# Per-phase host timing of the k-eigenvalue cycles on P ranks through the torch-free CLI launch.
# Usage: bash scripts/gpu_ktrace.sh P HISTORIES
P=${1:-2}; N=${2:-67108864}
D=$(mktemp -d)
python - "$D" "$N" <<'PY'
import sys
from minimc_b200 import decks
open(sys.argv[1] + "/deck.xml", "w").write(decks.k_infinite(histories=int(sys.argv[2]), threads=1, inactive=2, active=4))
PY
for r in $(seq 0 $((P-1))); do
  MMC_K_TRACE=1 MMC_WORLD_SIZE=$P MMC_RANK=$r MMC_DEVICE=$r MMC_COMM_ID_FILE=$D/id NCCL_DEBUG=${NCCL_DEBUG:-WARN} \
    minimc_b200/runminimc_b200 $D/deck.xml > $D/out.$r 2> $D/err.$r &
done
wait
for r in $(seq 0 $((P-1))); do echo "--- rank $r"; grep -E "cycle|NCCL|error" $D/err.$r | head -40; grep "k-effective" $D/out.$r; done
