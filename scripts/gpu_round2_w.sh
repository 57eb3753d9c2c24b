#!/bin/bash
# 1-GPU call W: multi-row stages in the LayerNorm-backward ring kernel, bias-gradient column sums fused into th_apply: tests, timing, bench
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_glue_gpu.py tests/test_th_gemm_gpu.py tests/test_cait_gpu.py tests/test_model_gpu.py -m gpu -q -x --timeout=200 -p no:cacheprovider > gpurun_out/w_tests.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/w_tests.log
for r in 1 2 4 0; do
  VITK_LN_RING_ROWS=$r timeout 300 python scripts/bench_ln.py 25088 384 2>&1 | tail -1 | tee -a gpurun_out/w_bench_ln.txt
done
for r in 1 2 0; do
  VITK_LN_RING_ROWS=$r timeout 300 python scripts/bench_ln.py 25216 768 2>&1 | tail -1 | tee -a gpurun_out/w_bench_ln.txt
done
timeout 300 python scripts/step_breakdown.py cait_S24_224 128 > gpurun_out/w_breakdown_cait.txt 2>&1; head -14 gpurun_out/w_breakdown_cait.txt
timeout 600 python bench.py --workload cait_S24_224 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-families --no-e2e > gpurun_out/w_bench_cait.json 2> gpurun_out/w_bench_cait.err
echo "bench cait rc=$?"; head -c 230 gpurun_out/w_bench_cait.json; echo
