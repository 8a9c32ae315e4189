#!/bin/bash
# last check of the final tree: what the driver runs at round end, plus ncu captures of the two decode kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r23_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r23_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:attn_decode_kernel" -s 3 -c 1 -f -o gpurun_out/r23_attn_decode python tools/gpu_decode_breakdown.py 2 > gpurun_out/r23_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gemm_skinny_kernel<\(int\)1, \(int\)1" -s 7 -c 1 -f -o gpurun_out/r23_skinny_down python tools/gpu_decode_breakdown.py 2 > gpurun_out/r23_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 900 python bench.py > gpurun_out/r23_bench.json 2> gpurun_out/r23_bench.err; echo "bench rc=$?"; wc -l gpurun_out/r23_bench.json; head -c 400 gpurun_out/r23_bench.json
