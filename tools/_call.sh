set -x
export PYTHONUNBUFFERED=1
# --- attention with one MMA issuer per tile: correctness, perf, timeline
timeout 600 python -m pytest tests/test_gpu_attn_adversarial.py tests/test_gpu_kernels.py -x -q > gpurun_out/attn2i_tests.txt 2>&1; echo tests rc=$?
tail -8 gpurun_out/attn2i_tests.txt
timeout 300 python tools/gpu_perf_attn.py > gpurun_out/attn_perf_2issuers.txt 2>&1; echo rc=$?
cat gpurun_out/attn_perf_2issuers.txt
timeout 300 python tools/gpu_attn_trace.py run 4096 0 > gpurun_out/attn_trace_2issuers.txt 2>&1; echo rc=$?
cat gpurun_out/attn_trace_2issuers.txt
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_variants.py tests/test_gpu_fullsize.py -x -q > gpurun_out/model_tests.txt 2>&1; echo tests rc=$?
tail -5 gpurun_out/model_tests.txt
# --- decode after the attention merge change
timeout 600 python tools/gpu_decode_ablate.py 28 32 1245 quick > gpurun_out/decode_ablate_v3.txt 2>&1; echo rc=$?
cat gpurun_out/decode_ablate_v3.txt
