#!/bin/bash
# Builds a variant of libminimc_b200.so with extra nvcc defines for kernels.cu and event_loop.cu (development
# experiments only):
#   scripts/build_variant.sh NAME -DMMC_CE_BLOCKS_PER_SM=2 ...
# -> minimc_b200/csrc/build/variants/NAME.so   (travels to the GPU box; scripts/gpu_variants.sh swaps it in there)
set -e
cd "$(dirname "$0")/../minimc_b200/csrc"
NAME=$1; shift
make -s >/dev/null
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall"
nvcc $FLAGS "$@" -c kernels.cu -o build/variants/$NAME.kernels.o &
nvcc $FLAGS "$@" -Xptxas -v -c event_loop.cu -o build/variants/$NAME.event_loop.o 2>&1 | grep -A2 "event_\(flight\|tsl\)_kernel" | grep -E "Used|spill" || true
wait
# (the counter-RNG build of the kernels, the C ABI and the host are the base build's)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$NAME.so build/variants/$NAME.kernels.o build/variants/$NAME.event_loop.o build/kernels_ctr.o build/event_loop_ctr.o build/capi.o build/comm.o build/host_*.o -ldl
rm build/variants/$NAME.kernels.o build/variants/$NAME.event_loop.o
echo built build/variants/$NAME.so
