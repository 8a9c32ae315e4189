#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "decode or attn_varlen" > gpurun_out/r12_tests.log 2>&1
echo "attn tests rc=$?"; tail -5 gpurun_out/r12_tests.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "7b_layer_dims" > gpurun_out/r12_tests_7b.log 2>&1
echo "7b-dims test rc=$?"; tail -8 gpurun_out/r12_tests_7b.log
timeout 300 python tools/gpu_perf_decode_kernels.py 32 1245 > gpurun_out/r12_decode_kernels.txt 2>&1; cat gpurun_out/r12_decode_kernels.txt
timeout 300 python tools/gpu_perf_decode_kernels.py 8 4096 2>&1 | sed -n 2p
timeout 300 python tools/gpu_perf_decode_kernels.py 1 8192 2>&1 | sed -n 1,6p
timeout 600 python tools/gpu_perf_prefill_decode.py > gpurun_out/r12_prefill_decode.txt 2>&1; tail -3 gpurun_out/r12_prefill_decode.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,smsp__inst_executed.sum --clock-control none --kernel-name-base demangled -k "regex:attn_decode_kernel" -s 3 -c 1 python tools/gpu_decode_breakdown.py 2 2>&1 | grep -E "attn_decode|duration|dram__|inst_executed" | head
