"""GPU diagnostic for bagel_gemm_bf16 (run under gpurun; prints an error map when a case is wrong)."""
import sys, time
import torch
sys.path.insert(0, ".")
from bagel_b200 import ops

torch.manual_seed(0)
dev = "cuda"

def ref_mm(a, w):
    return a.float() @ w.float().t()

def check(name, out, ref, tol_ulp=2.0):
    out = out.float(); 
    err = (out - ref).abs()
    tol = ref.abs() * (2.0 ** -8) * tol_ulp + 2e-3
    bad = (err > tol)
    nbad = int(bad.sum())
    print(f"[{name}] shape={tuple(out.shape)} max_abs_err={err.max().item():.4e} bad={nbad}/{out.numel()} "
          f"finite={bool(torch.isfinite(out).all())}", flush=True)
    if nbad:
        M, N = out.shape
        bm, bn = 32, 32
        mm = (M + bm - 1) // bm; nn = (N + bn - 1) // bn
        print("  error map (rows=32-row blocks, cols=32-col blocks; '#' = has bad element), first 16x16 blocks")
        for i in range(min(mm, 16)):
            line = ""
            for j in range(min(nn, 16)):
                blk = bad[i*bm:(i+1)*bm, j*bn:(j+1)*bn]
                line += "#" if blk.any() else "."
            print("  " + line)
        idx = bad.nonzero()[:8]
        for r, c in idx.tolist():
            print(f"   ({r},{c}) got {out[r,c].item():.5f} want {ref[r,c].item():.5f}")
    return nbad == 0

ok = True
cases = [(128, 256, 64), (128, 256, 256), (256, 512, 512), (384, 256, 3584), (300, 264, 4304), (128, 64, 64),
         (100, 128, 192), (4098, 4608, 3584), (16, 3584, 3584)]
for (M, N, K) in cases:
    a = (torch.randn(M, K, device=dev) ).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev).to(torch.bfloat16)
    try:
        out = ops.gemm(a, w, bias=bias)
        torch.cuda.synchronize()
    except Exception as e:
        print(f"[gemm {M}x{N}x{K}] EXCEPTION {e}", flush=True); ok = False; break
    ok &= check(f"bias {M}x{N}x{K}", out, ref_mm(a, w) + bias.float())

# residual epilogue
M, N, K = 512, 3584, 512
a = torch.randn(M, K, device=dev).to(torch.bfloat16); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
res = torch.randn(M, N, device=dev).to(torch.bfloat16)
out = ops.gemm(a, w, resid=res, epilogue=ops.EPI_RESID); torch.cuda.synchronize()
mm = ref_mm(a, w)
ref = res.float() + mm.to(torch.bfloat16).float()
# tolerance is relative to the operands (res, mm), not to the possibly-cancelling sum
err = (out.float() - ref).abs(); tol = (res.float().abs() + mm.abs()) * 2.0 ** -8 * 2 + 2e-3
print(f"[resid] max_abs_err={err.max().item():.4e} bad={int((err > tol).sum())}/{out.numel()}", flush=True)
ok &= bool((err <= tol).all())

# swiglu epilogue
M, I, K = 512, 1024, 256
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
gw = (torch.randn(I, K, device=dev) / K ** 0.5).to(torch.bfloat16); uw = (torch.randn(I, K, device=dev) / K ** 0.5).to(torch.bfloat16)
wi = ops.interleave_gate_up(gw, uw)
out = ops.gemm(a, wi, epilogue=ops.EPI_SWIGLU); torch.cuda.synchronize()
g = ref_mm(a, gw).to(torch.bfloat16); u = ref_mm(a, uw).to(torch.bfloat16)
ref = (torch.nn.functional.silu(g) * u).float()
ok &= check("swiglu", out, ref, tol_ulp=4.0)

# row_map scatter + gelu/silu
M, N, K = 64, 256, 128
a = torch.randn(M, K, device=dev).to(torch.bfloat16); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
big = torch.zeros(1000, N, device=dev, dtype=torch.bfloat16)
rm = (torch.randperm(1000, device=dev)[:M]).to(torch.int32)
ops.gemm(a, w, row_map=rm, out=big); torch.cuda.synchronize()
ok &= check("row_map", big[rm.long()], ref_mm(a, w))
out = ops.gemm(a, w, epilogue=ops.EPI_GELU); torch.cuda.synchronize()
ok &= check("gelu", out, torch.nn.functional.gelu(ref_mm(a, w).to(torch.bfloat16).float(), approximate="tanh"), tol_ulp=4.0)
out = ops.gemm(a, w, epilogue=ops.EPI_SILU); torch.cuda.synchronize()
ok &= check("silu", out, torch.nn.functional.silu(ref_mm(a, w).to(torch.bfloat16).float()), tol_ulp=4.0)

# timing of the model shapes (B=2 T2I rows) vs cuBLAS
def bench(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

if ok:
    for (M, N, K, epi) in [(8196, 4608, 3584, 0), (8196, 3584, 3584, 0), (8196, 37888, 3584, 2), (8196, 3584, 18944, 0),
                           (32784, 37888, 3584, 2), (32784, 3584, 18944, 0), (8192, 8192, 8192, 0)]:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
        out = torch.empty(M, N // 2 if epi == 2 else N, device=dev, dtype=torch.bfloat16)
        t = bench(lambda: ops.gemm(a, w, epilogue=epi, out=out))
        tc = bench(lambda: torch.matmul(a, w.t()))
        fl = 2.0 * M * N * K
        print(f"[perf] M={M} N={N} K={K} epi={epi}: ours {t:.3f} ms = {fl/t/1e9:.0f} TFLOP/s | cuBLAS {tc:.3f} ms = {fl/tc/1e9:.0f} TFLOP/s", flush=True)
print("ALL_OK" if ok else "SOME_FAILED")
