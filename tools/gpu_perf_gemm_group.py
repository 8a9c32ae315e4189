"""A/B of the GEMM raster group size on one box (each value needs a fresh process: the env var is read once)."""
import os, subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, ".")
from bagel_b200 import ops
def bench(fn, iters=6):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for (M, N, K, epi) in [(65568, 37888, 3584, 2), (65568, 3584, 18944, 0), (65568, 4608, 3584, 0), (65568, 3584, 3584, 0)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    out = torch.empty(M, N // 2 if epi == 2 else N, device="cuda", dtype=torch.bfloat16)
    t = bench(lambda: ops.gemm(a, w, epilogue=epi, out=out))
    print(f"  M={M} N={N} K={K} epi={epi}: {t:.3f} ms = {2.0*M*N*K/t/1e9:.0f} TFLOP/s", flush=True)
'''
for g in ("8", "16", "32", "64", "0"):
    env = dict(os.environ, BAGEL_GEMM_GROUP_M=g)
    print(f"group_m={g} (0 = auto)", flush=True)
    subprocess.run([sys.executable, "-c", code], env=env)
