set -x
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
# 1. the whole GPU suite
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_final.txt 2>&1; echo gpu tests rc=$?
tail -6 gpurun_out/gputest_final.txt
# 2. smoke
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo smoke rc=$?
tail -3 gpurun_out/smoke.txt
# 3. the bench line (all blocks)
timeout 1200 python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; echo bench rc=$?
tail -c 600 gpurun_out/bench_n1_final.err
# 4. reference arm
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?
# 5. ncu launch lists: one denoising step, one decode step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_step.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-extra --no-cpu-baseline --no-taylorseer > gpurun_out/ncu_step.log 2>&1; echo rc=$?
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/decode_launches_final.csv python tools/gpu_decode_breakdown.py 2 > gpurun_out/decode_breakdown_final.log 2>&1; echo rc=$?
# 6. ncu --set full of the dominant kernels
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm2 -s 1 -c 1 -f -o gpurun_out/ncu_gemm2_swiglu python tools/gpu_ncu_targets.py gemm > gpurun_out/ncu_gemm.log 2>&1; echo rc=$?
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn -s 1 -c 1 -f -o gpurun_out/ncu_attn_final python tools/gpu_ncu_targets.py attn > gpurun_out/ncu_attn.log 2>&1; echo rc=$?
ls -la gpurun_out/
