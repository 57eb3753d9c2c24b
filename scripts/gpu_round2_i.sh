#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -s --timeout=600 -p no:cacheprovider -k "split" > gpurun_out/i_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/i_tests.log; grep -E "dist parity|MISMATCH|passed|failed|rc=" gpurun_out/i_tests.log | tail -8
