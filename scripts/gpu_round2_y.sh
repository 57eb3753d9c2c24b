#!/bin/bash
# 1-GPU call Y: attention forward with two score tiles in flight and staggered softmax groups: parity, stagger sweep, fused backward at N = 785
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_attn_gpu.py tests/test_model_gpu.py -m gpu -q -x --timeout=200 -p no:cacheprovider > gpurun_out/y_tests.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/y_tests.log
for st in 0 200 350 500 800; do
  echo "stagger=$st"; VITK_ATTN_STAGGER=$st BENCH_ATTN_FWD_ONLY=1 timeout 200 python scripts/bench_attn.py gpurun_out/y_attn_fwd_st$st.json 2>&1 | grep -o "'shape': '[a-z0-9_]*'\|'fwd_us': [0-9.]*" | paste - - 
done
timeout 300 python scripts/bench_attn.py gpurun_out/y_bench_attn.json 2>&1 | tail -3
