set -x
export PYTHONUNBUFFERED=1
for v in halfrow halfrow2; do
BAGEL_TEST_LIB=tools/_trace/libbagel_b200_$v.so timeout 300 python -m pytest tests/test_gpu_attn_adversarial.py tests/test_gpu_kernels.py -x -q -k "attn or attention" -p no:cacheprovider > gpurun_out/attn_${v}_tests.txt 2>&1; echo $v tests rc=$?
tail -2 gpurun_out/attn_${v}_tests.txt
PERF_NO_FA2=1 PERF_LIB=tools/_trace/libbagel_b200_$v.so timeout 120 python tools/gpu_perf_attn.py > gpurun_out/attn_perf_$v.txt 2>&1; echo rc=$?
done
PERF_NO_FA2=1 timeout 120 python tools/gpu_perf_attn.py > gpurun_out/attn_perf_default.txt 2>&1; echo rc=$?
