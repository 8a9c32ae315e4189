"""A/B of the wide-GEMM raster (N super-tiles), TMA L2 hints and streaming output stores at the denoise shapes, one
process per setting (the env knobs are read once): BAGEL_GEMM_GROUP_N x BAGEL_GEMM_HINTS [x BAGEL_GEMM_GROUP_M]."""
import os, subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, ".")
from bagel_b200 import ops
def bench(fn, iters=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
res = []
for (M, N, K, epi) in [(65568, 37888, 3584, 2), (65568, 3584, 18944, 1)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    out = torch.empty(M, N // 2 if epi == 2 else N, device="cuda", dtype=torch.bfloat16)
    r = torch.randn(M, N, device="cuda").to(torch.bfloat16) if epi == 1 else None
    t = bench(lambda: ops.gemm(a, w, epilogue=epi, resid=r, out=out))
    res.append(f"N={N} K={K}: {t:.3f} ms {2.0*M*N*K/t/1e9:.0f} TF/s")
print("   " + " | ".join(res), flush=True)
'''
# (pair kernel on/off, group_n, hints, group_m in M-tiles)
cfgs = [(0, 0, 0, 0), (1, 0, 0, 0), (1, 0, 0, 24), (1, 0, 0, 32), (1, 0, 0, 48), (1, 0, 0, 64), (1, 0, 4, 16), (1, 0, 4, 32),
        (1, 0, 4, 48), (1, 37, 5, 32), (1, 0, 6, 32), (0, 0, 0, 0), (1, 0, 0, 32)]
for pair, gn, h, gm in cfgs:
    env = dict(os.environ, BAGEL_GEMM_PAIR=str(pair), BAGEL_GEMM_GROUP_N=str(gn), BAGEL_GEMM_HINTS=str(h),
               BAGEL_GEMM_GROUP_M=str(gm))
    print(f"pair={pair} group_n={gn} hints={h} group_m={gm or 'auto'}", flush=True)
    subprocess.run([sys.executable, "-c", code], env=env)
