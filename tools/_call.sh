set -x
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_attn_adversarial.py tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_train_forward.py -x -q -p no:cacheprovider > gpurun_out/attn_chunkmask_tests.txt 2>&1; echo tests rc=$?
tail -3 gpurun_out/attn_chunkmask_tests.txt
PERF_NO_FA2=1 PERF_LIB=tools/_trace/libbagel_b200_nochunkmask.so timeout 120 python tools/gpu_perf_attn.py > gpurun_out/attn_perf_nochunkmask.txt 2>&1; echo rc=$?
PERF_NO_FA2=1 timeout 120 python tools/gpu_perf_attn.py > gpurun_out/attn_perf_chunkmask.txt 2>&1; echo rc=$?
