#!/bin/bash
# A visit to a box with N GPUs (gpurun --gpus N): the multi-GPU parity tests, then bench.py under torchrun.
# Usage: bash scripts/gpu_multi.sh TAG N [stage ...]   stages: test bench keig strong
set -u
TAG=$1; N=$2; shift 2
STAGES=${*:-test bench keig strong}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for st in $STAGES; do case $st in
test) echo "== multi-GPU parity tests"; timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_multi.txt;;
bench) echo "== bench $N GPUs"; timeout 1500 $RUN bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2>$OUT/bench_${N}gpu.err | tee $OUT/bench_${N}gpu.json | cut -c1-300;;
keig) for wl in keigenvalue_mg keigenvalue_ce; do
  echo "== bench $wl $N GPUs"; timeout 900 $RUN bench.py --gpus $N --workload $wl --steps 8 --warmup 3 2>$OUT/bench_${wl}_${N}gpu.err | tee $OUT/bench_${wl}_${N}gpu.json | cut -c1-300; done;;
strong) echo "== bench strong scaling 1e8 over $N GPUs"; timeout 900 $RUN bench.py --gpus $N --scaling strong --total-histories 100000000 --steps 5 --warmup 3 --no-multigroup --no-cpu-baseline --no-extras 2>$OUT/bench_strong_${N}gpu.err | tee $OUT/bench_strong_${N}gpu.json | cut -c1-300;;
ktrace) echo "== k cycles, per-phase trace"; MMC_K_TRACE=1 timeout 600 $RUN bench.py --gpus $N --workload keigenvalue_mg --steps 4 --warmup 2 > $OUT/ktrace_${N}gpu.log 2>&1; tail -40 $OUT/ktrace_${N}gpu.log | cut -c1-200;;
*) echo "unknown stage $st";;
esac; done
ls -la $OUT
