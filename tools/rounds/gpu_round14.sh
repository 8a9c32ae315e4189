#!/bin/bash
for i in 1 2; do
echo "--- PDL on"; timeout 600 python tools/gpu_perf_prefill_decode.py 2>&1 | tail -1
echo "--- PDL off"; BAGEL_PDL=0 timeout 600 python tools/gpu_perf_prefill_decode.py 2>&1 | tail -1
done
echo "--- PDL off, skinny sweep shapes"; 
for shape in "32 3584 18944 1" "32 3584 3584 1" "32 4608 3584 0" "32 37888 3584 2"; do BAGEL_PDL=0 timeout 120 python tools/gpu_sweep_skinny.py $shape; timeout 120 python tools/gpu_sweep_skinny.py $shape; done
