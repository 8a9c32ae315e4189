set -x
timeout 600 python tools/gpu_decode_ablate.py > gpurun_out/decode_ablate.txt 2>&1; echo rc=$?
cat gpurun_out/decode_ablate.txt | tail -30
timeout 300 python tools/gpu_attn_trace.py run 4096 0 > gpurun_out/attn_trace_4096.txt 2>&1; echo rc=$?
timeout 300 python tools/gpu_attn_trace.py run 1024 1 16 > gpurun_out/attn_trace_1024c.txt 2>&1; echo rc=$?
cat gpurun_out/attn_trace_4096.txt gpurun_out/attn_trace_1024c.txt
