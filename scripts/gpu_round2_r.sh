#!/bin/bash
# 1-GPU call R: restructured mixing kernels (tests, per-op timing, bench) + ncu --set full of the fc1+GELU GEMM (ViT-B, CaiT)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_th_gemm_gpu.py tests/test_cait_gpu.py tests/test_golden.py -m gpu -q -x --timeout=120 -p no:cacheprovider > gpurun_out/r_tests.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r_tests.log
timeout 300 python scripts/step_breakdown.py cait_S24_224 128 > gpurun_out/r_breakdown_cait.txt 2>&1; head -12 gpurun_out/r_breakdown_cait.txt
timeout 600 python bench.py --workload cait_S24_224 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-families --no-e2e > gpurun_out/r_bench_cait.json 2> gpurun_out/r_bench_cait.err
echo "bench rc=$?"; head -c 230 gpurun_out/r_bench_cait.json; echo; tail -2 gpurun_out/r_bench_cait.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm -s 2 -c 1 -o gpurun_out/r_gemm_gelu_vitb -f python scripts/prof_gemm.py gelu 3072 768 > /dev/null 2>&1; echo "ncu gelu vitb rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm -s 2 -c 1 -o gpurun_out/r_gemm_gelu_cait -f python scripts/prof_gemm.py gelu 1536 384 > /dev/null 2>&1; echo "ncu gelu cait rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm -s 2 -c 1 -o gpurun_out/r_gemm_dgelu_vitb -f python scripts/prof_gemm.py dgelu 768 3072 > /dev/null 2>&1; echo "ncu dgelu vitb rc=$?"
ls -la gpurun_out | head -20
