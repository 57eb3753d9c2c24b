#!/bin/bash
# 8-GPU call V: deferred all-reduce with the arena registered with NCCL (NVLS zero-copy) vs plain vs bf16 wire; 1-GPU line of the same box
mkdir -p gpurun_out
for mode in registered deferred bf16; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 20 --warmup 5 --dp $mode --no-families --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/v_bench_8gpu_$mode.json 2> gpurun_out/v_bench_8gpu_$mode.err
  echo "bench 8gpu $mode rc=$?"; head -c 200 gpurun_out/v_bench_8gpu_$mode.json; echo; grep -E "registered|Error" gpurun_out/v_bench_8gpu_$mode.err | tail -2
done
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-families --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/v_bench_1gpu.json 2> gpurun_out/v_bench_1gpu.err
echo "bench 1gpu rc=$?"; head -c 200 gpurun_out/v_bench_1gpu.json; echo
