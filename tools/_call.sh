set -x
export PYTHONUNBUFFERED=1
# --- V3c: Q tile in TMEM (tcgen05.cp), single O buffer
BAGEL_ATTN_V3=1 BAGEL_ATTN_QT=1 timeout 600 python -m pytest tests/test_gpu_attn_adversarial.py tests/test_gpu_kernels.py -x -q -k "attn or attention" > gpurun_out/v3c_tests.txt 2>&1; echo v3c tests rc=$?
tail -15 gpurun_out/v3c_tests.txt
BAGEL_ATTN_V3=1 BAGEL_ATTN_QT=1 PERF_NO_FA2=1 timeout 300 python tools/gpu_perf_attn.py > gpurun_out/attn_perf_v3c.txt 2>&1; echo rc=$?
cat gpurun_out/attn_perf_v3c.txt
# --- ncu full of the attention kernels (stall reasons)
BAGEL_ATTN_V3=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn -s 1 -c 1 -f -o gpurun_out/ncu_attn_v3b python tools/gpu_ncu_targets.py attn > gpurun_out/ncu_v3b.log 2>&1; echo rc=$?
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn -s 1 -c 1 -f -o gpurun_out/ncu_attn_v2 python tools/gpu_ncu_targets.py attn > gpurun_out/ncu_v2.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
