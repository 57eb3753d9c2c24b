#!/bin/bash
# Round 2, GPU call A: full GPU test-suite, then bench lines of the four GPU workloads (no profiler).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/a_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_tests.log
tail -40 gpurun_out/a_tests.log
for wl in dino_vitb16 cait_S24_224 dino_vitb16_lineareval dino_vitb8; do
  extra="--no-cpu-baseline"
  [ "$wl" = "dino_vitb16" ] && extra=""
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 $extra > gpurun_out/a_bench_$wl.json 2> gpurun_out/a_bench_$wl.err
  echo "bench $wl rc=$?"; head -c 600 gpurun_out/a_bench_$wl.json; echo; tail -3 gpurun_out/a_bench_$wl.err
done
