#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:gemm_bf16_kernel|gemm_skinny_kernel|attn_varlen_kernel|attn_decode_kernel|rmsnorm_kernel|qk_norm_rope_kernel|copy_rows_kernel|rope_table_kernel|latent_embed_add_kernel|cfg_norm_kernel|cfg_apply_kernel|cast_f32_bf16_kernel" -c 3000 --csv --log-file gpurun_out/r20_launches_bench_step.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r20_ncu_list.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/r20_launches_bench_step.csv
for g in 0 32 24 48; do
  echo "=== BAGEL_GEMM_GROUP_M=$g (0 = built-in heuristic 16 wide / 32 narrow)"
  BAGEL_GEMM_GROUP_M=$g timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 8 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['achieved'], d['clocks']['sm_mhz'])"
done
