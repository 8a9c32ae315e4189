#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "attn" > gpurun_out/r16_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r16_tests.log
timeout 300 python tools/gpu_perf_attn_tiles.py
timeout 600 python tools/gpu_perf_attn.py 2>&1 | grep -E "ours" | cut -c1-120
