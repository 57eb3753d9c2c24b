#!/bin/bash
# 1-GPU call M: new short-sequence talking-heads products: unit tests, CaiT parity tests, CaiT bench + per-op breakdown
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_th_gemm_gpu.py -m gpu -q -x --timeout=120 -p no:cacheprovider > gpurun_out/m_tests_thgemm.log 2>&1
echo "pytest th_gemm rc=$?"; tail -15 gpurun_out/m_tests_thgemm.log
timeout 600 python -m pytest tests/test_cait_gpu.py tests/test_golden.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/m_tests_cait.log 2>&1
echo "pytest cait rc=$?"; tail -8 gpurun_out/m_tests_cait.log
timeout 300 python scripts/step_breakdown.py cait_S24_224 128 > gpurun_out/m_breakdown_cait.txt 2>&1; head -24 gpurun_out/m_breakdown_cait.txt
timeout 600 python bench.py --workload cait_S24_224 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-families > gpurun_out/m_bench_cait.json 2> gpurun_out/m_bench_cait.err
echo "bench rc=$?"; head -c 300 gpurun_out/m_bench_cait.json; echo; tail -3 gpurun_out/m_bench_cait.err
