#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
cp xmem2_b200/libxmem2_b200.so /tmp/lib_backup.so
XMEM_EXTRA_NVCC_FLAGS=-DK1_TRACE python -c "from xmem2_b200 import build; build.build(force=True)" > gpurun_out/r2n_build.log 2>&1; echo "build rc=$?"
K1_TRACE=1 timeout 120 python tests/profile_k1_timeline.py > gpurun_out/r2n_trace.txt 2>&1; tail -45 gpurun_out/r2n_trace.txt
cp /tmp/lib_backup.so xmem2_b200/libxmem2_b200.so
