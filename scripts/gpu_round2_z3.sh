#!/bin/bash
# 1-GPU call Z3 (validated attention forward): MMA-warp polling with / without back-off -- parity, timing, trace with MMA-side events
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_attn_gpu.py -m gpu -q -x -k "test_attn_fwd" --timeout=30 -p no:cacheprovider > gpurun_out/z3_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/z3_tests.log
for sp in 0 1; do
  echo "mma_spin=$sp"; VITK_ATTN_MMA_SPIN=$sp BENCH_ATTN_FWD_ONLY=1 timeout 60 python scripts/bench_attn.py gpurun_out/z3_attn_fwd_spin$sp.json 2>&1 | grep -o "'shape': '[a-z0-9_]*'\|'fwd_us': [0-9.]*" | paste - -
done
export VITK_LIB=$PWD/vit_torch_b200/libvitk_dbg.so
for sp in 0 1; do
  VITK_ATTN_MMA_SPIN=$sp timeout 60 python scripts/trace_attn.py fwd 128 197 12 64 > gpurun_out/z3_trace_fwd_vitb16_spin$sp.txt 2>&1
  echo "trace spin=$sp"; grep -E "lifetime|first scores|tile [12]: elementwise \(|tile [12]: done|tile [12]: elementwise done|MMA warp sees|last tile|write-out" gpurun_out/z3_trace_fwd_vitb16_spin$sp.txt | grep -v "tile [03]:"
done
