#!/bin/bash
# GEMM / LayerNorm part of the profile pass (run under gpurun, one GPU): per-launch DRAM traffic of every GEMM launch
# of one step, and --set full captures of the fc1+GELU GEMM, the qkv forward GEMM and the LayerNorm backward.
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm \
    -s 438 -c 146 --csv --log-file gpurun_out/r01_gemm_traffic.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 2 -c 1 -o gpurun_out/r01_gemm2_fc1_gelu -f \
    python scripts/prof_gemm.py gelu 3072 768 > /dev/null 2>&1; echo "ncu full gelu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 2 -c 1 -o gpurun_out/r01_gemm2_qkv_fwd -f \
    python scripts/prof_gemm.py fwd 2304 768 > /dev/null 2>&1; echo "ncu full fwd rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ln_bwd -s 1 -c 1 -o gpurun_out/r01_ln_bwd -f \
    python scripts/prof_attn.py ln > /dev/null 2>&1; echo "ncu full ln rc=$?"
