#!/bin/bash
# split-K / pipeline-depth sweep of the skinny GEMM at the decode shapes (B=32)
mkdir -p gpurun_out
O=gpurun_out/r9_skinny_sweep.txt; : > $O
for shape in "32 3584 18944 1" "32 3584 3584 1" "32 4608 3584 0"; do
  for sp in 1 2 4 5 7 8; do
    BAGEL_SKINNY_SPLIT=$sp timeout 120 python tools/gpu_sweep_skinny.py $shape >> $O 2>&1
  done
  for st in 4 6 8; do
    BAGEL_SKINNY_STAGES=$st timeout 120 python tools/gpu_sweep_skinny.py $shape >> $O 2>&1
  done
  timeout 120 python tools/gpu_sweep_skinny.py $shape >> $O 2>&1
done
for st in 3 4 6; do BAGEL_SKINNY_STAGES=$st timeout 120 python tools/gpu_sweep_skinny.py 32 37888 3584 2 >> $O 2>&1; done
timeout 120 python tools/gpu_sweep_skinny.py 32 37888 3584 2 >> $O 2>&1
timeout 120 python tools/gpu_sweep_skinny.py 32 152064 3584 0 >> $O 2>&1
BAGEL_GEMM_SKINNY=0 timeout 120 python tools/gpu_sweep_skinny.py 32 3584 18944 1 >> $O 2>&1
cat $O
