#!/bin/bash
# decode-path round: new kernels' parity tests, then microbench + end-to-end decode timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "skinny or decode or argmax or attn" > gpurun_out/r5_tests_kernels.log 2>&1
echo "kernel tests rc=$?"; tail -5 gpurun_out/r5_tests_kernels.log
timeout 300 python tools/gpu_perf_decode_kernels.py 32 1245 > gpurun_out/r5_decode_kernels.txt 2>&1; cat gpurun_out/r5_decode_kernels.txt
timeout 900 python -m pytest tests/test_gpu_model.py -x -q > gpurun_out/r5_tests_model.log 2>&1
echo "model tests rc=$?"; tail -5 gpurun_out/r5_tests_model.log
timeout 600 python tools/gpu_perf_prefill_decode.py > gpurun_out/r5_prefill_decode.txt 2>&1; tail -4 gpurun_out/r5_prefill_decode.txt
