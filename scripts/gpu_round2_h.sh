#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -q -s --timeout=600 -p no:cacheprovider -k "split or deferred" > gpurun_out/h_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_tests.log; grep -E "dist parity|MISMATCH|passed|failed|rc=|Error|error" gpurun_out/h_tests.log | tail -20
for mode in split deferred; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --dp $mode --no-families --no-e2e > gpurun_out/h_bench_2gpu_$mode.json 2> gpurun_out/h_bench_2gpu_$mode.err
  echo "bench 2gpu $mode rc=$?"; head -c 250 gpurun_out/h_bench_2gpu_$mode.json; echo; tail -3 gpurun_out/h_bench_2gpu_$mode.err
done
