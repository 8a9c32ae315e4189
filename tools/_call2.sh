set -x
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err; echo rc=$?
tail -c 800 gpurun_out/bench_n2_final.err
head -c 600 gpurun_out/bench_n2_final.json
