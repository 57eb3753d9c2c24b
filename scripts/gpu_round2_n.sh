#!/bin/bash
# 1-GPU call N: A/B of the talking-heads kernel variants (per-op timing inside a CaiT-S24 step), parity tests per variant,
# ncu launch list of one CaiT block
mkdir -p gpurun_out
for cfg in "2 2" "1 1" "0 2"; do
  set -- $cfg
  export VITK_TH_BWD=$1 VITK_TH_APPLY=$2
  timeout 300 python -m pytest tests/test_th_gemm_gpu.py tests/test_cait_gpu.py -m gpu -q -x --timeout=120 -p no:cacheprovider > gpurun_out/n_tests_$1$2.log 2>&1
  echo "variant bwd=$1 apply=$2: pytest rc=$?"; tail -3 gpurun_out/n_tests_$1$2.log
  timeout 300 python scripts/step_breakdown.py cait_S24_224 128 > gpurun_out/n_breakdown_cait_$1$2.txt 2>&1; head -9 gpurun_out/n_breakdown_cait_$1$2.txt
done
unset VITK_TH_BWD VITK_TH_APPLY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/n_launches_cait_block.csv \
    python scripts/prof_cait_block.py > gpurun_out/n_ncu_block.log 2>&1; echo "ncu block launches rc=$?"
ls -la gpurun_out | head -30
