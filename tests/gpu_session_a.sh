#!/bin/bash
# round-2 GPU session A: full GPU suite, bench line, probes, experimental convolution variants (each in its own process)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
python -m pytest tests -m gpu -q --timeout 900 -rA 2>&1 | tail -80 > gpurun_out/r2a_gputest.txt
tail -5 gpurun_out/r2a_gputest.txt
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.json
python tests/bench_conv.py > gpurun_out/r2a_bench_conv_plain.txt 2>&1; tail -1 gpurun_out/r2a_bench_conv_plain.txt
(cd tests/microbench && timeout 60 ./umma_view_probe > ../../gpurun_out/r2a_umma_probe.txt 2>&1; timeout 60 ./pair_mma_probe > ../../gpurun_out/r2a_pair_probe.txt 2>&1; timeout 200 ./tma_ingest > ../../gpurun_out/r2a_tma_ingest.txt 2>&1)
for v in halo csk 2cta mc; do
  XMEM_CONV_IMPL=$v timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 120 2>&1 | tail -15 > gpurun_out/r2a_conv_$v.txt
  tail -2 gpurun_out/r2a_conv_$v.txt
  XMEM_CONV_IMPL=$v timeout 300 python tests/bench_conv.py > gpurun_out/r2a_bench_conv_$v.txt 2>&1; tail -1 gpurun_out/r2a_bench_conv_$v.txt
done
XMEM_RUN_UNVERIFIED=1 timeout 200 python -m pytest tests/test_gpu_zz_selector.py -m gpu -q --timeout 120 2>&1 | tail -15 > gpurun_out/r2a_selector_fused.txt
tail -2 gpurun_out/r2a_selector_fused.txt
