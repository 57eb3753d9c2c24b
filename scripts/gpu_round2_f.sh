#!/bin/bash
# 2-GPU call: NCCL gradient parity + DP bench modes
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_dist_gpu.py -m gpu -q -s --timeout=900 -p no:cacheprovider > gpurun_out/f_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_tests.log; grep -E "dist parity|MISMATCH|passed|failed|rc=|Error" gpurun_out/f_tests.log | tail -20
for mode in deferred bf16; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --dp $mode --no-families > gpurun_out/f_bench_2gpu_$mode.json 2> gpurun_out/f_bench_2gpu_$mode.err
  echo "bench 2gpu $mode rc=$?"; head -c 250 gpurun_out/f_bench_2gpu_$mode.json; echo; tail -2 gpurun_out/f_bench_2gpu_$mode.err
done
