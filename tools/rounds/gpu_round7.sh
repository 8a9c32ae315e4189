#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "taylor" > gpurun_out/r7_tests_taylor_kernels.log 2>&1
echo "taylor kernel tests rc=$?"; tail -5 gpurun_out/r7_tests_taylor_kernels.log
timeout 900 python -m pytest tests/test_gpu_model.py -x -q > gpurun_out/r7_tests_model.log 2>&1
echo "model tests rc=$?"; tail -8 gpurun_out/r7_tests_model.log
timeout 900 python bench.py > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r7_bench.json; tail -3 gpurun_out/r7_bench.err
