#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit=$?"; tail -12 gpurun_out/pytest_gpu.txt
echo "== attention perf"; timeout 300 python tools/gpu_perf_attn.py > gpurun_out/perf_attn.txt 2>&1; tail -14 gpurun_out/perf_attn.txt
echo "== bench"; timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
echo "== aux perf"; timeout 900 python tools/gpu_perf_aux.py > gpurun_out/perf_aux.txt 2>&1; grep -v "^\s" gpurun_out/perf_aux.txt | tail -12
