#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_perf_decode_kernels.py 32 1245 > gpurun_out/r6_decode_kernels.txt 2>&1; cat gpurun_out/r6_decode_kernels.txt
timeout 300 python tools/gpu_perf_decode_kernels.py 8 4096 > gpurun_out/r6_decode_kernels_b8.txt 2>&1; tail -9 gpurun_out/r6_decode_kernels_b8.txt
timeout 600 python tools/gpu_perf_prefill_decode.py > gpurun_out/r6_prefill_decode.txt 2>&1; tail -5 gpurun_out/r6_prefill_decode.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:gemm_|attn_|rmsnorm_kernel|qk_norm_rope_kernel|copy_rows_kernel|rope_table_kernel|decode_|argmax_rows" -c 400 --csv --log-file gpurun_out/r6_decode_launches.csv python tools/gpu_decode_breakdown.py 2 > gpurun_out/r6_decode_ncu.log 2>&1; tail -2 gpurun_out/r6_decode_ncu.log
