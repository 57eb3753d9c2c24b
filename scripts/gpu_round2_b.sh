#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/b_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_tests.log
tail -25 gpurun_out/b_tests.log
for wl in cait_S24_224 dino_vitb8; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/b_bench_$wl.json 2> gpurun_out/b_bench_$wl.err
  echo "bench $wl rc=$?"; head -c 300 gpurun_out/b_bench_$wl.json; echo; tail -3 gpurun_out/b_bench_$wl.err
done
timeout 600 python scripts/step_breakdown.py cait_S24_224 128 > gpurun_out/b_breakdown_cait.txt 2>&1; tail -30 gpurun_out/b_breakdown_cait.txt
