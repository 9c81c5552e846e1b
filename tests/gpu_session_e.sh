#!/bin/bash
# K1 iteration: parity tests, phase timeline, roofline at both shapes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_k1.py -x -q > gpurun_out/r2g_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2g_pytest.txt
timeout 120 python tests/profile_k1_timeline.py > gpurun_out/r2g_timeline.txt 2>&1; cat gpurun_out/r2g_timeline.txt | tail -24
timeout 120 python tests/profile_k1_timeline.py 8160 6 5 1 > gpurun_out/r2g_timeline_1080.txt 2>&1; tail -22 gpurun_out/r2g_timeline_1080.txt
timeout 200 python - <<'PY' 2>&1 | tee gpurun_out/r2g_roofline.txt
import json, bench, torch
torch.set_grad_enabled(False)
r = bench.k1_roofline('cuda:0'); print(r['launch_us'], r['frac'])
r = bench.k1_roofline('cuda:0', hw=8160, frames=(6, 5)); print(r['launch_us'], r['frac'])
PY
