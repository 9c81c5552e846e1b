#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_k1.py tests/test_gpu_clip.py -x -q > gpurun_out/r2j_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2j_pytest.txt
timeout 120 python tests/profile_k1_timeline.py > gpurun_out/r2j_timeline.txt 2>&1; head -22 gpurun_out/r2j_timeline.txt | grep -v "MMA phase\|slowest"
timeout 300 python tests/diag_k1_in_clip.py 98 > gpurun_out/r2j_k1_in_clip.txt 2>&1; tail -9 gpurun_out/r2j_k1_in_clip.txt
timeout 600 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1])
print(l['value'], l['e2e']['value'], l['roofline']['launch_us'], l['roofline']['frac'], l['roofline_1080p_shape']['launch_us'], l['roofline_1080p_shape']['frac'], l.get('check'), l.get('speedup_vs_reference_style_gpu'))
PY
