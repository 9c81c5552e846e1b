#!/bin/bash
# round-2 evidence session: tests, bench, conv/eltwise ncu captures, racecheck
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.txt
timeout 600 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; head -c 1500 gpurun_out/r2f_bench.json
python -m xmem2_b200.util.conv_bench > gpurun_out/r2f_conv_table.txt 2>&1; tail -1 gpurun_out/r2f_conv_table.txt
i=0
for shape in "fuser 3x3 1600" "up8 3x3 256" "fuser 3x3 512->512" "l3 3x3 256" "up16 3x3 512->256" "keyproj"; do
  i=$((i+1))
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_ --launch-skip 3 -c 1 -f -o gpurun_out/r2f_conv_$i \
      python -m xmem2_b200.util.conv_bench "$shape" > gpurun_out/r2f_ncu_conv_$i.log 2>&1; echo "ncu conv $i rc=$?"
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"im2col_stem3|cbam_|conv3x3_c1|upsample|maxpool|area_down" --profile-from-start off --launch-skip 30 -c 16 -f -o gpurun_out/r2f_eltwise \
    python tests/profile_clip.py 12 > gpurun_out/r2f_ncu_eltwise.log 2>&1; echo "ncu eltwise rc=$?"
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_conv.py -x -q > gpurun_out/r2f_racecheck_conv.txt 2>&1; echo "racecheck conv rc=$?"; tail -4 gpurun_out/r2f_racecheck_conv.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_k1.py -x -q -k "single_bank_small or three_banks or two_objects" > gpurun_out/r2f_racecheck_k1.txt 2>&1; echo "racecheck k1 rc=$?"; tail -4 gpurun_out/r2f_racecheck_k1.txt
