#!/bin/bash
mkdir -p gpurun_out
echo "== pytest (graph + model)"; timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu > gpurun_out/pytest_model.txt 2>&1; echo "pytest exit=$?"; tail -6 gpurun_out/pytest_model.txt
echo "== bench (graphs in e2e)"; timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; tail -3 gpurun_out/bench_graph.err; python -c "
import json;d=json.load(open('gpurun_out/bench_graph.json'));print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])"
echo "== compute-sanitizer memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/sanitizer_memcheck.txt 2>&1; echo "memcheck exit=$?"; tail -8 gpurun_out/sanitizer_memcheck.txt
echo "== compute-sanitizer racecheck"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_small.py > gpurun_out/sanitizer_racecheck.txt 2>&1; echo "racecheck exit=$?"; tail -8 gpurun_out/sanitizer_racecheck.txt
