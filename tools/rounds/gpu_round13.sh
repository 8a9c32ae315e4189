#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "skinny or decode" > gpurun_out/r13_tests.log 2>&1
echo "kernel tests rc=$?"; tail -4 gpurun_out/r13_tests.log
timeout 900 python -m pytest tests/test_gpu_model.py -x -q > gpurun_out/r13_tests_model.log 2>&1
echo "model tests rc=$?"; tail -4 gpurun_out/r13_tests_model.log
echo "--- PDL on"; timeout 600 python tools/gpu_perf_prefill_decode.py 2>&1 | tail -2
echo "--- PDL off"; BAGEL_PDL=0 timeout 600 python tools/gpu_perf_prefill_decode.py 2>&1 | tail -2
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/r13_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r13_sanitizer_memcheck.txt
