#!/bin/bash
set -u
mkdir -p gpurun_out
export VITK_LIB=$PWD/vit_torch_b200/libvitk_dbg.so
for w in ${WHICH:-fwd dq dkv}; do timeout 120 python scripts/trace_attn.py $w 2>&1 | tee gpurun_out/trace_$w.txt; done
