#!/bin/bash
# 1-GPU call X (end-of-round validation): smoke, full GPU suite, the four bench lines with baselines and family rooflines,
# ncu launch lists of the headline and the CaiT command
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/x_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/x_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/x_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/x_tests.log; tail -6 gpurun_out/x_tests.log
for wl in dino_vitb16 cait_S24_224 dino_vitb16_lineareval dino_vitb8; do
  extra="--no-cpu-baseline"
  [ "$wl" = "dino_vitb16" ] && extra=""
  timeout 900 python bench.py --workload $wl --steps 20 --warmup 5 $extra > gpurun_out/x_bench_$wl.json 2> gpurun_out/x_bench_$wl.err
  echo "bench $wl rc=$?"; head -c 300 gpurun_out/x_bench_$wl.json; echo; tail -2 gpurun_out/x_bench_$wl.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/x_launches_vitb16.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-families > gpurun_out/x_ncu_vitb16.log 2>&1; echo "ncu vitb16 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/x_launches_cait.csv \
    python bench.py --workload cait_S24_224 --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-families > gpurun_out/x_ncu_cait.log 2>&1; echo "ncu cait rc=$?"
ls -la gpurun_out | head -30
