#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 > gpurun_out/r2_final_bench_4gpu.json 2> gpurun_out/r2_final_bench_4gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_final_bench_4gpu.json').read().strip().splitlines()[-1])
print(l['value'], l['e2e']['value'], json.dumps(l.get('tshard'))[:900])
PY
