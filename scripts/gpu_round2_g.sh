#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cait_gpu.py tests/test_regressions_gpu.py tests/test_checkpoint.py -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/g_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/g_tests.log; tail -6 gpurun_out/g_tests.log
timeout 600 python scripts/step_breakdown.py cait_S24_224 128 > gpurun_out/g_breakdown_cait.txt 2>&1; head -14 gpurun_out/g_breakdown_cait.txt
timeout 900 python bench.py --workload cait_S24_224 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/g_bench_cait.json 2> gpurun_out/g_bench_cait.err
echo "bench rc=$?"; head -c 200 gpurun_out/g_bench_cait.json; echo
