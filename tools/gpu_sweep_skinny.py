"""Time ONE skinny-GEMM shape under CUDA-graph replay (20 launches cycling over 4 weight copies, so every launch
streams its weights from HBM and no host launch overhead is measured). Knobs come from the environment
(BAGEL_SKINNY_SPLIT / BAGEL_SKINNY_STAGES are read once per process), hence one process per configuration.
Usage: python tools/gpu_sweep_skinny.py M N K epi(0 bias|1 resid|2 swiglu)"""
import os, sys
import torch
sys.path.insert(0, ".")
from bagel_b200 import ops

M, N, K, epi = (int(a) for a in sys.argv[1:5])
dev, bf = "cuda", torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
ws = [(torch.randn(N, K, device=dev, generator=g) * K ** -0.5).to(bf) for _ in range(4)]
a = torch.randn(M, K, device=dev, generator=g).to(bf)
n_out = N // 2 if epi == 2 else N
res = torch.randn(M, n_out, device=dev, generator=g).to(bf)
out = torch.empty(M, n_out, device=dev, dtype=bf)


def run(i):
    if epi == 1:
        ops.gemm(a, ws[i % 4], resid=res, epilogue=ops.EPI_RESID, out=out)
    elif epi == 2:
        ops.gemm(a, ws[i % 4], epilogue=ops.EPI_SWIGLU, out=out)
    else:
        ops.gemm(a, ws[i % 4], out=out)


for i in range(4):
    run(i)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    for i in range(20):
        run(i)
graph.replay(); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3 / 20)
t = sorted(ts)[len(ts) // 2]
nbytes = N * K * 2
print(f"M={M} N={N} K={K} epi={epi} split={os.environ.get('BAGEL_SKINNY_SPLIT', 'auto')} "
      f"stages={os.environ.get('BAGEL_SKINNY_STAGES', 'auto')}: {t:7.2f} us/launch  {nbytes / t / 1e6:5.2f} TB/s", flush=True)
