#!/bin/bash
# K1 bring-up: parity tests of the fused kernel, then the clips, then the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_k1.py -x -q --timeout 180 2>&1 | tail -40 > gpurun_out/r2b_k1.txt
tail -15 gpurun_out/r2b_k1.txt
if grep -q "passed" gpurun_out/r2b_k1.txt && ! grep -q "failed" gpurun_out/r2b_k1.txt; then
  timeout 900 python -m pytest tests/test_gpu_tshard.py tests/test_gpu_clip.py tests/test_gpu_zz_chair.py tests/test_gpu_zz_selector.py -q --timeout 300 2>&1 | tail -30 > gpurun_out/r2b_rest.txt
  tail -8 gpurun_out/r2b_rest.txt
  timeout 600 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 900 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
fi
