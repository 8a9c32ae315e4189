#!/bin/bash
mkdir -p gpurun_out
for sp in 2 1; do
  echo "=== BAGEL_ATTN_SPLIT=$sp"
  BAGEL_ATTN_SPLIT=$sp timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "attn" > gpurun_out/r15_tests_split$sp.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r15_tests_split$sp.log
  BAGEL_ATTN_SPLIT=$sp timeout 300 python tools/gpu_perf_attn_tiles.py
  BAGEL_ATTN_SPLIT=$sp timeout 600 python tools/gpu_perf_attn.py 2>&1 | grep -E "ours" | cut -c1-120
done
