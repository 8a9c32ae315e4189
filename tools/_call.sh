set -x
timeout 1200 python bench.py > gpurun_out/bench_n1_final2.json 2> gpurun_out/bench_n1_final2.err; echo bench rc=$?
tail -c 300 gpurun_out/bench_n1_final2.err
