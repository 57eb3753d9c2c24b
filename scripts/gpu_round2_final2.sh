#!/bin/bash
# final 2-GPU sanity of the committed tree: NCCL gradient parity (all modes), headline line in the default (auto) mode
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_dist_gpu.py -m gpu -q -s --timeout=200 -p no:cacheprovider > gpurun_out/f2_tests.log 2>&1
echo "pytest rc=$?"; grep -E "dist parity|MISMATCH|passed|failed|Error" gpurun_out/f2_tests.log | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/f2_bench_2gpu.json 2> gpurun_out/f2_bench_2gpu.err
echo "bench 2gpu rc=$?"; head -c 230 gpurun_out/f2_bench_2gpu.json; echo; tail -2 gpurun_out/f2_bench_2gpu.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/f2_ref_2gpu.json 2> gpurun_out/f2_ref_2gpu.err
echo "reference arm rc=$?"; head -c 300 gpurun_out/f2_ref_2gpu.json; echo
