#!/bin/bash
# 1-GPU call O: th_scores with TMA-store epilogue: tests, per-op timing, then ncu --set full of the three talking-heads
# kernel families (one launch each, no source import: small reports)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_th_gemm_gpu.py tests/test_cait_gpu.py -m gpu -q -x --timeout=120 -p no:cacheprovider > gpurun_out/o_tests.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/o_tests.log
timeout 300 python scripts/step_breakdown.py cait_S24_224 128 > gpurun_out/o_breakdown_cait.txt 2>&1; head -10 gpurun_out/o_breakdown_cait.txt
for k in th_scores th_apply th_mix2_bwd th_mix2_fwd; do
  timeout 300 ncu --set full --clock-control none -k regex:$k -s 1 -c 1 -o gpurun_out/o_$k -f python scripts/prof_cait_block.py > /dev/null 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out
