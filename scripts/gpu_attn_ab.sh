#!/bin/bash
# Attention kernels: parity + microbench (VITK_ATTN_WG2=0/1 selects the one-/two-group kernels). Run under gpurun.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attn_gpu.py -x -q --timeout 120 > gpurun_out/pytest_attn.log 2>&1; echo "pytest attn rc=$?"
tail -3 gpurun_out/pytest_attn.log
for v in ${VARIANTS:-1}; do
  VITK_ATTN_WG2=$v timeout 300 python scripts/bench_attn.py > gpurun_out/bench_attn_wg$v.log 2>&1; echo "wg$v rc=$?"
  grep "'op'" gpurun_out/bench_attn_wg$v.log
done
if [ "${BENCH:-0}" = "1" ]; then
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "bench rc=$?"
  cut -c1-200 gpurun_out/bench_ab.json
fi
