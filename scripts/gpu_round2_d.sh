#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attn_gpu.py -m gpu -q --timeout=300 -p no:cacheprovider -k "bwd" > gpurun_out/d_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_tests.log; tail -15 gpurun_out/d_tests.log
timeout 600 python scripts/bench_attn.py gpurun_out/d_bench_attn.json 2>&1 | tail -5
