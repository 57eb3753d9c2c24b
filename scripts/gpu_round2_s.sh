#!/bin/bash
# 1-GPU call S: back-off polling in the GEMM producer / MMA roles (A/B: VITK_GEMM_DBG=16 = bare polling), CaiT parity after the mixing-kernel edits
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_th_gemm_gpu.py tests/test_cait_gpu.py -m gpu -q -x --timeout=120 -p no:cacheprovider > gpurun_out/s_tests.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/s_tests.log
for dbg in 0 16 0 16; do
  VITK_GEMM_DBG=$dbg timeout 300 python scripts/bench_epi.py 2>&1 | tail -1 | tee -a gpurun_out/s_bench_epi.txt
done
for dbg in 0 16; do
  VITK_GEMM_DBG=$dbg timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-families --no-e2e > gpurun_out/s_bench_vitb16_dbg$dbg.json 2> gpurun_out/s_bench_vitb16_dbg$dbg.err
  echo "bench vitb16 dbg=$dbg rc=$?"; head -c 230 gpurun_out/s_bench_vitb16_dbg$dbg.json; echo
done
timeout 600 python bench.py --workload cait_S24_224 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-families --no-e2e > gpurun_out/s_bench_cait.json 2> gpurun_out/s_bench_cait.err
echo "bench cait rc=$?"; head -c 230 gpurun_out/s_bench_cait.json; echo
