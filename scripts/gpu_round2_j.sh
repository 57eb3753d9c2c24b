#!/bin/bash
# Round 2, GPU call J (re-entry): full GPU suite, the four bench lines, ncu launch list of the headline command,
# attention micro-bench, per-op step breakdowns. Everything lands in gpurun_out/j_*.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/j_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j_tests.log
tail -8 gpurun_out/j_tests.log
for wl in dino_vitb16 cait_S24_224 dino_vitb16_lineareval dino_vitb8; do
  extra="--no-cpu-baseline"
  [ "$wl" = "dino_vitb16" ] && extra=""
  timeout 900 python bench.py --workload $wl --steps 20 --warmup 5 $extra > gpurun_out/j_bench_$wl.json 2> gpurun_out/j_bench_$wl.err
  echo "bench $wl rc=$?"; head -c 400 gpurun_out/j_bench_$wl.json; echo; tail -3 gpurun_out/j_bench_$wl.err
done
timeout 600 python scripts/bench_attn.py gpurun_out/j_bench_attn.json 2>&1 | tail -5
timeout 600 python scripts/step_breakdown.py dino_vitb16 128 > gpurun_out/j_breakdown_vitb16.txt 2>&1; head -30 gpurun_out/j_breakdown_vitb16.txt
timeout 600 python scripts/step_breakdown.py cait_S24_224 128 > gpurun_out/j_breakdown_cait.txt 2>&1; head -30 gpurun_out/j_breakdown_cait.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/j_launches_vitb16.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-families > gpurun_out/j_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
