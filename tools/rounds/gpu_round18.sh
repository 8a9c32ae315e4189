#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attn" > gpurun_out/r18_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r18_tests.log
timeout 300 python tools/gpu_perf_attn_tiles.py
timeout 600 python tools/gpu_perf_attn.py 2>&1 | grep -E "ours" | cut -c1-125
echo "=== BAGEL_ATTN_HALVES=0"
BAGEL_ATTN_HALVES=0 timeout 300 python tools/gpu_perf_attn_tiles.py
