set -x
export PYTHONUNBUFFERED=1
# --- attention V3 (double-buffered S): correctness, then perf vs the two-tile kernel on the same box
BAGEL_ATTN_V3=1 timeout 600 python -m pytest tests/test_gpu_attn_adversarial.py tests/test_gpu_kernels.py -x -q -k "attn or attention" > gpurun_out/v3_tests.txt 2>&1; echo v3 tests rc=$?
tail -15 gpurun_out/v3_tests.txt
BAGEL_ATTN_V3=1 timeout 300 python tools/gpu_perf_attn.py > gpurun_out/attn_perf_v3.txt 2>&1; echo rc=$?
timeout 300 python tools/gpu_perf_attn.py > gpurun_out/attn_perf_v2.txt 2>&1; echo rc=$?
cat gpurun_out/attn_perf_v3.txt gpurun_out/attn_perf_v2.txt
# --- decode: in-graph ablation, knobs
BAGEL_QKROPE_SPREAD=0 timeout 600 python tools/gpu_decode_ablate.py 28 32 1245 > gpurun_out/decode_ablate_base.txt 2>&1; echo rc=$?
timeout 300 python tools/gpu_decode_ablate.py 28 32 1245 quick > gpurun_out/decode_ablate_spread.txt 2>&1; echo rc=$?
BAGEL_PDL_SMALL=1 timeout 300 python tools/gpu_decode_ablate.py 28 32 1245 quick > gpurun_out/decode_ablate_pdlsmall.txt 2>&1; echo rc=$?
BAGEL_PDL_SMALL=1 BAGEL_DECODE_SPLIT=4 timeout 300 python tools/gpu_decode_ablate.py 28 32 1245 quick > gpurun_out/decode_ablate_pdlsmall_split4.txt 2>&1; echo rc=$?
BAGEL_PDL_SMALL=1 BAGEL_DECODE_SPLIT=1 timeout 300 python tools/gpu_decode_ablate.py 28 32 1245 quick > gpurun_out/decode_ablate_pdlsmall_split1.txt 2>&1; echo rc=$?
tail -n 30 gpurun_out/decode_ablate_*.txt
# --- timeline of the two-tile attention kernel
timeout 300 python tools/gpu_attn_trace.py run 4096 0 > gpurun_out/attn_trace_4096.txt 2>&1; echo rc=$?
cat gpurun_out/attn_trace_4096.txt
# --- VAE after the gn_finalize restructure
timeout 600 python -m pytest tests/test_gpu_vae.py -x -q > gpurun_out/vae_tests.txt 2>&1; echo vae tests rc=$?
tail -3 gpurun_out/vae_tests.txt
timeout 600 python tools/gpu_perf_aux.py > gpurun_out/vae_siglip_timing.txt 2>&1; echo rc=$?
tail -8 gpurun_out/vae_siglip_timing.txt
