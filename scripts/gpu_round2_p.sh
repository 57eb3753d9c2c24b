#!/bin/bash
# 1-GPU call P: side-stream overlap of the independent talking-heads products: parity tests, CaiT bench with and without
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cait_gpu.py tests/test_train_gpu.py -m gpu -q -x --timeout=200 -p no:cacheprovider > gpurun_out/p_tests.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/p_tests.log
for ss in 1 0; do
  VITK_SIDE_STREAM=$ss timeout 600 python bench.py --workload cait_S24_224 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-families --no-e2e > gpurun_out/p_bench_cait_ss$ss.json 2> gpurun_out/p_bench_cait_ss$ss.err
  echo "bench side_stream=$ss rc=$?"; head -c 230 gpurun_out/p_bench_cait_ss$ss.json; echo; tail -2 gpurun_out/p_bench_cait_ss$ss.err
done
