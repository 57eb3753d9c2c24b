#!/bin/bash
# 1-GPU call Q: persistent th_scores / th_apply: tests, per-op timing (heads per item 2 vs 4), CaiT bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_th_gemm_gpu.py tests/test_cait_gpu.py -m gpu -q -x --timeout=120 -p no:cacheprovider > gpurun_out/q_tests.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/q_tests.log
for hg in 0 4 1; do
  VITK_TH_APPLY_HG=$hg timeout 300 python scripts/step_breakdown.py cait_S24_224 128 > gpurun_out/q_breakdown_cait_hg$hg.txt 2>&1; echo "hg=$hg"; head -9 gpurun_out/q_breakdown_cait_hg$hg.txt
done
timeout 600 python bench.py --workload cait_S24_224 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-families --no-e2e > gpurun_out/q_bench_cait.json 2> gpurun_out/q_bench_cait.err
echo "bench rc=$?"; head -c 230 gpurun_out/q_bench_cait.json; echo; tail -2 gpurun_out/q_bench_cait.err
