set -x
export PYTHONUNBUFFERED=1
timeout 300 python tools/gpu_attn_trace.py run 4096 0 8 > gpurun_out/attn_trace_fine.txt 2>&1; echo rc=$?
cat gpurun_out/attn_trace_fine.txt
BAGEL_ATTN_POLY=0 timeout 300 python tools/gpu_attn_trace.py run 4096 0 8 > gpurun_out/attn_trace_fine_nopoly.txt 2>&1; echo rc=$?
cat gpurun_out/attn_trace_fine_nopoly.txt
