set -x
export PYTHONUNBUFFERED=1
# --- attention: 5 K/V stages (default now) vs 4 (variant lib), poly exp2 on/off; correctness of both new paths
timeout 600 python -m pytest tests/test_gpu_attn_adversarial.py tests/test_gpu_kernels.py -x -q -k "attn or attention" > gpurun_out/attn_s5_tests.txt 2>&1; echo tests rc=$?
tail -3 gpurun_out/attn_s5_tests.txt
BAGEL_ATTN_POLY=1 timeout 600 python -m pytest tests/test_gpu_attn_adversarial.py tests/test_gpu_kernels.py -x -q -k "attn or attention" > gpurun_out/attn_poly_tests.txt 2>&1; echo tests rc=$?
tail -3 gpurun_out/attn_poly_tests.txt
PERF_NO_FA2=1 PERF_LIB=tools/_trace/libbagel_b200_stages4.so timeout 300 python tools/gpu_perf_attn.py > gpurun_out/attn_perf_s4.txt 2>&1; echo rc=$?
PERF_NO_FA2=1 timeout 300 python tools/gpu_perf_attn.py > gpurun_out/attn_perf_s5.txt 2>&1; echo rc=$?
PERF_NO_FA2=1 BAGEL_ATTN_POLY=1 timeout 300 python tools/gpu_perf_attn.py > gpurun_out/attn_perf_s5_poly.txt 2>&1; echo rc=$?
PERF_NO_FA2=1 BAGEL_ATTN_POLY=1 PERF_LIB=tools/_trace/libbagel_b200_stages4.so timeout 300 python tools/gpu_perf_attn.py > gpurun_out/attn_perf_s4_poly.txt 2>&1; echo rc=$?
tail -n 14 gpurun_out/attn_perf_s4.txt gpurun_out/attn_perf_s5.txt gpurun_out/attn_perf_s5_poly.txt gpurun_out/attn_perf_s4_poly.txt
