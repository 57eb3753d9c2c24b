#!/bin/bash
# Round profile pass (run under gpurun, one GPU): bench line, ncu launch list of the same command, and one --set full
# capture of each attention kernel. (GEMM traffic / GEMM and LayerNorm captures: scripts/gpu_profile_gemm.sh.)
set -u
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
python scripts/prof_attn.py > /dev/null 2>&1; echo "plain attn rc=$?"
for k in attn_fwd2 attn_bwd_dq2 attn_bwd_dkv2; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/r01_$k -f \
      python scripts/prof_attn.py > /dev/null 2>&1; echo "ncu full $k rc=$?"
done
