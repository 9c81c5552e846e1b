#!/bin/sh
# Build the standalone probes in-tree (the binaries are git-ignored but travel with gpurun), then e.g.
#   gpurun --timeout 300 -- 'cd tests/microbench && timeout 60 ./umma_view_probe; timeout 60 ./pair_mma_probe; timeout 120 ./tma_ingest'
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Wno-deprecated-gpu-targets"
$NVCC $FLAGS -o tma_ingest tma_ingest.cu -lcuda
$NVCC $FLAGS -I../../xmem2_b200/csrc -o umma_view_probe umma_view_probe.cu ../../xmem2_b200/csrc/common.cu
$NVCC $FLAGS -I../../xmem2_b200/csrc -o pair_mma_probe pair_mma_probe.cu ../../xmem2_b200/csrc/common.cu
ls -la tma_ingest umma_view_probe pair_mma_probe
