#!/bin/bash
# final 1-GPU sanity of the committed tree: smoke, full GPU suite, headline line (no baselines)
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/f1_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f1_smoke.log
timeout 600 python -m pytest tests -m gpu -q --timeout=120 -p no:cacheprovider > gpurun_out/f1_tests.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/f1_tests.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-families > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err
echo "bench rc=$?"; head -c 230 gpurun_out/f1_bench.json; echo
