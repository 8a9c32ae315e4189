#!/bin/bash
mkdir -p gpurun_out
for cfg in "BAGEL_ATTN_POLY=0" "BAGEL_ATTN_POLY=8" "BAGEL_ATTN_POLY=4" "BAGEL_ATTN_POLY=3" "BAGEL_ATTN_POLY=2" "BAGEL_ATTN_SPLIT=2"; do
  echo "=== $cfg"
  env $cfg timeout 300 python tools/gpu_perf_attn_tiles.py
  env $cfg timeout 600 python tools/gpu_perf_attn.py 2>&1 | grep -E "denoise|L=16384 nseq=1 causal=False|H=28/4 L=1024 nseq=16 causal=False" | cut -c1-110
done
BAGEL_ATTN_POLY=4 timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "attn" 2>&1 | tail -2
