set -x
export PYTHONUNBUFFERED=1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/sanitizer_memcheck.txt 2>&1; echo rc=$?
tail -3 gpurun_out/sanitizer_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/sanitizer_racecheck.txt 2>&1; echo rc=$?
tail -3 gpurun_out/sanitizer_racecheck.txt
BAGEL_ATTN_V3=1 timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/sanitizer_memcheck_v3.txt 2>&1; echo rc=$?
tail -3 gpurun_out/sanitizer_memcheck_v3.txt
