#!/bin/bash
# 1-GPU call L: pipe-rate micro-benchmark, CaiT launch list, --set full capture of one CaiT block (fwd + bwd kernels)
mkdir -p gpurun_out
scripts/micro/bin/pipe_rates > gpurun_out/l_pipe_rates.txt 2>&1; cat gpurun_out/l_pipe_rates.txt
python scripts/prof_cait_block.py > /dev/null 2>&1; echo "plain block rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -s 34 -c 34 -o gpurun_out/l_cait_block -f \
    python scripts/prof_cait_block.py > gpurun_out/l_ncu_block.log 2>&1; echo "ncu block rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/l_launches_cait.csv \
    python bench.py --workload cait_S24_224 --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-families > gpurun_out/l_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
