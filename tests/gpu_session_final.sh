#!/bin/bash
# round-2 closing evidence: full GPU suite, K1 ncu capture -> traffic json, launch list, bench (both arms)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
REV="$1"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_final_pytest.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_fused --launch-skip 1 -c 1 -f -o gpurun_out/r2_k1_fused python tests/profile_k1.py 2 > gpurun_out/r2_final_ncu_k1.log 2>&1; echo "ncu k1 rc=$?"
ncu -i gpurun_out/r2_k1_fused.ncu-rep --page raw --csv > gpurun_out/r2_k1_raw.csv 2>/dev/null
python tests/k1_traffic_json.py gpurun_out/r2_k1_raw.csv profiles/r2_k1_traffic.json "git $REV, ncu --set full --clock-control none, profiles/r2_k1_fused.ncu-rep" && cp profiles/r2_k1_traffic.json gpurun_out/r2_k1_traffic.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_clip_eager.csv python tests/profile_clip.py 40 > gpurun_out/r2_final_ncu_list.log 2>&1; echo "ncu list rc=$?"
python tests/summarize_launches.py gpurun_out/r2_launches_clip_eager.csv > gpurun_out/r2_launches_clip_eager.summary.txt 2>&1; head -12 gpurun_out/r2_launches_clip_eager.summary.txt
gzip -f gpurun_out/r2_launches_clip_eager.csv
timeout 200 python tests/profile_k1_timeline.py 8160 6 5 1 > gpurun_out/r2_k1_timeline_1080p.txt 2>&1
timeout 600 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; echo "bench ref rc=$?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1])
print(l['value'], l['e2e']['value'], l['roofline'], l['roofline_1080p_shape']['launch_us'], l['roofline_1080p_shape']['frac'], l['conv_roofline']['frame_us'], l.get('check'), l.get('speedup_vs_reference_style_gpu'), l['cpu_baseline'], l['gpu_launches'])
print(open('gpurun_out/r2_final_bench_reference.json').read()[:600])
PY
