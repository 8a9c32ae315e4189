#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "decode or attn_varlen" > gpurun_out/r8_tests_attn.log 2>&1
echo "attn tests rc=$?"; tail -5 gpurun_out/r8_tests_attn.log
timeout 300 python tools/gpu_perf_decode_kernels.py 32 1245 > gpurun_out/r8_decode_kernels.txt 2>&1; cat gpurun_out/r8_decode_kernels.txt
timeout 300 python tools/gpu_perf_decode_kernels.py 8 4096 > gpurun_out/r8_decode_kernels_b8.txt 2>&1; sed -n 2p gpurun_out/r8_decode_kernels_b8.txt
timeout 600 python tools/gpu_perf_prefill_decode.py > gpurun_out/r8_prefill_decode.txt 2>&1; tail -3 gpurun_out/r8_prefill_decode.txt
