#!/bin/bash
# 2-GPU call U: gradient arena registered with NCCL (ncclMemAlloc + register_mem_pool): parity, then bench next to deferred
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_dist_gpu.py -m gpu -q -s --timeout=300 -p no:cacheprovider -k "registered" > gpurun_out/u_tests.log 2>&1
echo "pytest rc=$?"; grep -E "dist parity|arena registered|MISMATCH|passed|failed|Error|not registered" gpurun_out/u_tests.log | tail -8
for mode in registered deferred; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 --dp $mode --no-families --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/u_bench_2gpu_$mode.json 2> gpurun_out/u_bench_2gpu_$mode.err
  echo "bench 2gpu $mode rc=$?"; head -c 200 gpurun_out/u_bench_2gpu_$mode.json; echo; grep -E "registered|Error" gpurun_out/u_bench_2gpu_$mode.err | tail -3
done
