"""Small single-purpose workloads for `ncu --set full` captures (one kernel family each; run under ncu with -k / -s / -c):
  python tools/gpu_ncu_targets.py gemm     gate|up + SwiGLU GEMM at the benchmark shape (M=65568, N=37888, K=3584), 3 launches
  python tools/gpu_ncu_targets.py down     down_proj + residual (M=65568, N=3584, K=18944), 3 launches
  python tools/gpu_ncu_targets.py attn     denoise-shaped attention (B=16, q=4098, kv=4164, GQA 28:4), 3 launches
  python tools/gpu_ncu_targets.py vae      one FLUX-VAE decode at 1024^2 (after a warm-up decode)
"""
import sys
import torch
sys.path.insert(0, ".")
from bagel_b200 import ops, synthetic

what = sys.argv[1]
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
if what in ("gemm", "down"):
    M, N, K, epi = (65568, 37888, 3584, ops.EPI_SWIGLU) if what == "gemm" else (65568, 3584, 18944, ops.EPI_RESID)
    a = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev, generator=g) / K ** 0.5).to(torch.bfloat16)
    out = torch.empty(M, N // 2 if epi == ops.EPI_SWIGLU else N, device=dev, dtype=torch.bfloat16)
    r = torch.randn(M, N, device=dev, generator=g).to(torch.bfloat16) if epi == ops.EPI_RESID else None
    for _ in range(3):
        ops.gemm(a, w, epilogue=epi, resid=r, out=out)
elif what == "attn":
    lq, lk = [4098] * 16, [4164] * 16
    q = torch.randn(sum(lq), 28, 128, device=dev, generator=g).to(torch.bfloat16)
    k = torch.randn(sum(lk), 4, 128, device=dev, generator=g).to(torch.bfloat16)
    v = torch.randn(sum(lk), 4, 128, device=dev, generator=g).to(torch.bfloat16)
    cq = torch.tensor([0] + torch.tensor(lq).cumsum(0).tolist(), dtype=torch.int32, device=dev)
    ck = torch.tensor([0] + torch.tensor(lk).cumsum(0).tolist(), dtype=torch.int32, device=dev)
    out = torch.empty_like(q)
    for _ in range(3):
        ops.attn_varlen(q, k, v, cq, ck, 4098, 4164, False, out=out)
elif what == "vae":
    vae = synthetic.build_random_vae(dev)
    z = torch.randn(1, 16, 128, 128, device=dev, generator=g)
    vae.decode(z)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("decode")
    vae.decode(z)
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("done", what)
