#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest "tests/test_model_gpu.py" "tests/test_cait_gpu.py::test_cait_matches_oracle" "tests/test_train_gpu.py::test_loss_curve_overlays_oracle_200_steps" -m gpu -q -s --timeout=600 -p no:cacheprovider > gpurun_out/e_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e_tests.log; grep -E "worst|nerr=|loss start|passed|failed|rc=" gpurun_out/e_tests.log | tail -30
