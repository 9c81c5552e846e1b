#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tshard.py tests/test_gpu_zz_tshard_clip.py -x -q > gpurun_out/r2_final_pytest_2gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_final_pytest_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 > gpurun_out/r2_final_bench_2gpu.json 2> gpurun_out/r2_final_bench_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_final_bench_2gpu.json').read().strip().splitlines()[-1])
print(l['value'], l['e2e']['value'], json.dumps(l.get('tshard'))[:900])
PY
