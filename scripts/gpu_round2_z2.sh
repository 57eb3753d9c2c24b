#!/bin/bash
# 1-GPU call Z2: attention forward with S issued two tiles ahead; MMA-warp polling with / without back-off: parity, timing, trace
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attn_gpu.py -m gpu -q -x --timeout=200 -p no:cacheprovider > gpurun_out/z2_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/z2_tests.log
for sp in 0 1; do
  echo "mma_spin=$sp"; VITK_ATTN_MMA_SPIN=$sp BENCH_ATTN_FWD_ONLY=1 timeout 200 python scripts/bench_attn.py gpurun_out/z2_attn_fwd_spin$sp.json 2>&1 | grep -o "'shape': '[a-z0-9_]*'\|'fwd_us': [0-9.]*" | paste - -
done
export VITK_LIB=$PWD/vit_torch_b200/libvitk_dbg.so
for sp in 0 1; do
  VITK_ATTN_MMA_SPIN=$sp timeout 120 python scripts/trace_attn.py fwd 128 197 12 64 > gpurun_out/z2_trace_fwd_vitb16_spin$sp.txt 2>&1
  echo "trace spin=$sp"; grep -E "lifetime|first scores|tile [12]: elementwise \(|tile [12]: done|tile [12]: elementwise done|last tile|write-out" gpurun_out/z2_trace_fwd_vitb16_spin$sp.txt
done
