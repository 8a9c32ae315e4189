#!/bin/bash
# final round-1 validation: full GPU suite, attention microbench + ncu capture of the shipped attention kernel inside the
# bench step, launch list of one bench step, N=1 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r19_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r19_pytest_gpu.log
timeout 600 python tools/gpu_perf_attn.py > gpurun_out/r19_perf_attn.txt 2>&1; tail -3 gpurun_out/r19_perf_attn.txt | cut -c1-120
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:attn_varlen_kernel' -s 2 -c 1 -f -o gpurun_out/r19_attn python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --layers 2 > gpurun_out/r19_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r19_launches_bench_step.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r19_ncu_list.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/r19_launches_bench_step.csv
timeout 900 python bench.py > gpurun_out/r19_bench.json 2> gpurun_out/r19_bench.err; echo "bench rc=$?"; tail -c 1200 gpurun_out/r19_bench.json
