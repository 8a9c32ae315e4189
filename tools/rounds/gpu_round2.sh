#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/pytest_gpu.txt
tail -25 gpurun_out/pytest_gpu.txt
echo "== ncu full: swiglu gemm"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm_bf16_kernel<\(int\)256, \(int\)2>' -s 2 -c 1 -f -o gpurun_out/prof_gemm_swiglu python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --layers 2 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit=$?"
tail -5 gpurun_out/ncu_gemm.log | cut -c1-300
ls -la gpurun_out
