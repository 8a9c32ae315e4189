set -x
timeout 120 tools/_trace/tmem_mufu > gpurun_out/microbench_tmem_mufu.txt 2>&1; echo rc=$?
cat gpurun_out/microbench_tmem_mufu.txt
