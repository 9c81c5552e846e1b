#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_clip.py tests/test_gpu_zz_chair.py tests/test_gpu_baseline_shapes.py -x -q > gpurun_out/r2m_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2m_pytest.txt
XMEM_NO_PDL=1 timeout 300 python tests/profile_gaps.py 100 > gpurun_out/r2m_gaps_nopdl.txt 2>&1; echo "rc=$?"; head -4 gpurun_out/r2m_gaps_nopdl.txt | tail -2; grep -A200 "kernels from one memory read" gpurun_out/r2m_gaps_nopdl.txt | sed -n 1,2p
timeout 300 python tests/profile_gaps.py 100 > gpurun_out/r2m_gaps_pdl.txt 2>&1; head -4 gpurun_out/r2m_gaps_pdl.txt | tail -2
