#!/bin/bash
# 2-GPU call: NCCL gradient parity tests, then the headline bench in each data-parallel mode (same box, back to back).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -s --timeout=600 -p no:cacheprovider > gpurun_out/k_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/k_tests.log; grep -E "dist parity|MISMATCH|passed|failed|rc=|Error" gpurun_out/k_tests.log | tail -12
for mode in split deferred bf16 overlap; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --dp $mode --no-families --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/k_bench_2gpu_$mode.json 2> gpurun_out/k_bench_2gpu_$mode.err
  echo "bench 2gpu $mode rc=$?"; head -c 200 gpurun_out/k_bench_2gpu_$mode.json; echo; tail -2 gpurun_out/k_bench_2gpu_$mode.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-families --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/k_bench_1gpu.json 2> gpurun_out/k_bench_1gpu.err
echo "bench 1gpu rc=$?"; head -c 200 gpurun_out/k_bench_1gpu.json; echo
