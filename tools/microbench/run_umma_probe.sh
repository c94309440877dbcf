#!/bin/bash
# Build and run every tcgen05 probe, one process per probe (an illegal encoding kills only its own context).
#   gpurun --timeout 300 -- 'bash tools/microbench/run_umma_probe.sh > gpurun_out/umma_probe.txt 2>&1'
set -u
cd "$(dirname "$0")"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a umma_probe.cu -o umma_probe || exit 1
for p in $(./umma_probe list); do
    timeout 30 ./umma_probe "$p" || echo "$p: exit $?"
done
