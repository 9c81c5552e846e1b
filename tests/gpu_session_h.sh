#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_network.py tests/test_gpu_clip.py tests/test_gpu_baseline_shapes.py -x -q > gpurun_out/r2o_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_pytest.txt
python -m xmem2_b200.util.conv_bench > gpurun_out/r2o_conv_table.txt 2>&1; tail -12 gpurun_out/r2o_conv_table.txt
timeout 300 python tests/profile_gaps.py 100 > gpurun_out/r2o_gaps_pdl.txt 2>&1; head -4 gpurun_out/r2o_gaps_pdl.txt | tail -2
