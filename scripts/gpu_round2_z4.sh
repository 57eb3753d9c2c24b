#!/bin/bash
# 1-GPU call Z4: skew between the eight softmax warps of the attention forward (trace), ViT-B/16 shape; both row blocks apart
mkdir -p gpurun_out
export VITK_LIB=$PWD/vit_torch_b200/libvitk_dbg.so
timeout 60 python scripts/trace_attn.py fwd 128 197 12 64 > gpurun_out/z4_trace_fwd_vitb16.txt 2>&1
grep -E "lifetime|tile 1" gpurun_out/z4_trace_fwd_vitb16.txt
timeout 60 python scripts/trace_attn.py fwd 128 256 12 64 > gpurun_out/z4_trace_fwd_n256.txt 2>&1
echo "N=256 (no dead warps)"; grep -E "lifetime|tile 1" gpurun_out/z4_trace_fwd_n256.txt
