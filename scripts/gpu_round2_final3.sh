#!/bin/bash
# last sanity of the final tree (attention forward restored to the validated kernel): smoke, attention + model tests, headline line
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 80 python -m pytest tests/test_attn_gpu.py tests/test_model_gpu.py tests/test_train_gpu.py -m gpu -q -x --timeout=40 -p no:cacheprovider 2>&1 | tail -2
timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-families --no-e2e | head -c 200; echo
