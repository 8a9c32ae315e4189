#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "decode or attn_varlen or skinny" > gpurun_out/r10_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r10_tests.log
timeout 300 python tools/gpu_perf_decode_kernels.py 32 1245 > gpurun_out/r10_decode_kernels.txt 2>&1; cat gpurun_out/r10_decode_kernels.txt
timeout 300 python tools/gpu_perf_decode_kernels.py 8 4096 > gpurun_out/r10_decode_kernels_b8.txt 2>&1; sed -n 1,6p gpurun_out/r10_decode_kernels_b8.txt
for sp in 1 2 4 8; do BAGEL_DECODE_SPLIT=$sp timeout 120 python tools/gpu_perf_decode_kernels.py 32 1245 2>&1 | sed -n 2p; done
timeout 600 python tools/gpu_perf_prefill_decode.py > gpurun_out/r10_prefill_decode.txt 2>&1; tail -3 gpurun_out/r10_prefill_decode.txt
