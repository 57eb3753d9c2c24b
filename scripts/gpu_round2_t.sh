#!/bin/bash
# 8-GPU call T: headline bench in the default (split) and the deferred data-parallel mode, same box, back to back
mkdir -p gpurun_out
for mode in split deferred; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --dp $mode --no-families --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/t_bench_8gpu_$mode.json 2> gpurun_out/t_bench_8gpu_$mode.err
  echo "bench 8gpu $mode rc=$?"; head -c 230 gpurun_out/t_bench_8gpu_$mode.json; echo; tail -2 gpurun_out/t_bench_8gpu_$mode.err
done
