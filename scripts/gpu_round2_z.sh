#!/bin/bash
# 1-GPU call Z: per-CTA event trace of the attention forward (instrumented build), ViT-B/16 and ViT-B/8 shapes
mkdir -p gpurun_out
export VITK_LIB=$PWD/vit_torch_b200/libvitk_dbg.so
timeout 120 python scripts/trace_attn.py fwd 128 197 12 64 > gpurun_out/z_trace_fwd_vitb16.txt 2>&1; cat gpurun_out/z_trace_fwd_vitb16.txt
timeout 120 python scripts/trace_attn.py fwd 64 785 12 64 > gpurun_out/z_trace_fwd_vitb8.txt 2>&1; head -60 gpurun_out/z_trace_fwd_vitb8.txt
