#!/bin/bash
# full GPU suite, sanitizer on the new kernels, ncu captures of the decode kernels, final N=1 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r11_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r11_pytest_gpu.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/r11_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r11_sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/r11_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r11_sanitizer_racecheck.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_skinny_kernel<\(int\)2" -s 3 -c 1 -f -o gpurun_out/r11_skinny_swiglu python tools/gpu_decode_breakdown.py 2 > gpurun_out/r11_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:attn_decode_kernel" -s 3 -c 1 -f -o gpurun_out/r11_attn_decode python tools/gpu_decode_breakdown.py 2 > gpurun_out/r11_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 900 python bench.py > gpurun_out/r11_bench.json 2> gpurun_out/r11_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r11_bench.json
