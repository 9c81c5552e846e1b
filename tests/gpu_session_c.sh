#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_k1.py tests/test_gpu_tshard.py -x -q --timeout 180 2>&1 | tail -30 > gpurun_out/r2c_k1.txt
tail -12 gpurun_out/r2c_k1.txt
python tests/profile_k1_timeline.py 1620 9 5 1 2>&1 | tail -13
python tests/profile_k1_timeline.py 8160 5 1 1 2>&1 | tail -13
