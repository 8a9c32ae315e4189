#!/bin/bash
# One GPU-box visit: tests, smoke, bench, ncu launch list + full captures of the two top kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt
echo "== bench" ; timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" == "ncu" ]; then
KREGEX='regex:gemm_bf16_kernel|attn_varlen_kernel|rmsnorm_kernel|qk_norm_rope_kernel|copy_rows_kernel|latent_embed_add_kernel|cfg_norm_kernel|cfg_apply_kernel|cast_f32_bf16_kernel|rope_table_kernel'
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 1400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit=$?"
echo "== ncu full: swiglu gemm"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm_bf16_kernel<256, *2>' -s 30 -c 1 -f -o gpurun_out/prof_gemm_swiglu python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --layers 2 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit=$?"
echo "== ncu full: attention"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:attn_varlen_kernel' -s 3 -c 1 -f -o gpurun_out/prof_attn python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --layers 2 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit=$?"
ls -la gpurun_out
fi
