set -x
export PYTHONUNBUFFERED=1
# --- attention with elect.sync issuers: v2 (default), v3b, v3b+QT
timeout 600 python -m pytest tests/test_gpu_attn_adversarial.py tests/test_gpu_kernels.py -x -q > gpurun_out/attn_elect_tests.txt 2>&1; echo tests rc=$?
tail -5 gpurun_out/attn_elect_tests.txt
timeout 300 python tools/gpu_perf_attn.py > gpurun_out/attn_perf_v2_elect.txt 2>&1; echo rc=$?
cat gpurun_out/attn_perf_v2_elect.txt
BAGEL_ATTN_V3=1 PERF_NO_FA2=1 timeout 300 python tools/gpu_perf_attn.py > gpurun_out/attn_perf_v3b_elect.txt 2>&1; echo rc=$?
cat gpurun_out/attn_perf_v3b_elect.txt
BAGEL_ATTN_V3=1 BAGEL_ATTN_QT=1 PERF_NO_FA2=1 timeout 300 python tools/gpu_perf_attn.py > gpurun_out/attn_perf_v3c_elect.txt 2>&1; echo rc=$?
cat gpurun_out/attn_perf_v3c_elect.txt
timeout 300 python tools/gpu_attn_trace.py run 4096 0 > gpurun_out/attn_trace_elect.txt 2>&1; echo rc=$?
cat gpurun_out/attn_trace_elect.txt
# --- headline step with the elect.sync issuers in every tcgen05 kernel
timeout 600 python bench.py --no-extra --no-e2e --no-cpu-baseline --no-taylorseer --steps 8 --warmup 3 > gpurun_out/bench_quick_elect.json 2> gpurun_out/bench_quick_elect.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick_elect.json'))
print({k:d[k] for k in ('ms_per_step','value','clocks')}, d['roofline']['achieved'], d['roofline']['avg_launch_ms'])
PY
timeout 600 python tools/gpu_decode_ablate.py 28 32 1245 quick > gpurun_out/decode_ablate_v4.txt 2>&1; echo rc=$?
head -3 gpurun_out/decode_ablate_v4.txt
