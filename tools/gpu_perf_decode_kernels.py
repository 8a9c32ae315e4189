"""Per-kernel timing of one text-decode layer at 7B shapes (B tokens, ctx keys): skinny GEMMs, decode attention,
argmax. Weights are cycled over NSETS distinct copies (each layer set is 466 MB > L2) so every launch streams from HBM.
Usage: python tools/gpu_perf_decode_kernels.py [B] [ctx]"""
import sys
import torch
sys.path.insert(0, ".")
from bagel_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
CTX = int(sys.argv[2]) if len(sys.argv) > 2 else 1245
dev = "cuda"
H, I, HQ, HK, D, V = 3584, 18944, 28, 4, 128, 152064
NSETS = 4
bf = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(bf)


sets = [dict(wqkv=rnd(4608, H, scale=H ** -0.5), bqkv=rnd(4608), wo=rnd(H, H, scale=H ** -0.5),
             wgu=rnd(2 * I, H, scale=H ** -0.5), wd=rnd(H, I, scale=I ** -0.5)) for _ in range(NSETS)]
head = rnd(V, H, scale=H ** -0.5)
x = rnd(B, H)
act = rnd(B, I)
qkv = torch.empty(B, 4608, device=dev, dtype=bf)
out = torch.empty(B, H, device=dev, dtype=bf)
logits = torch.empty(B, V, device=dev, dtype=bf)
cap = CTX + 16
kbuf = [rnd(B * cap, HK, D) for _ in range(NSETS)]
vbuf = [rnd(B * cap, HK, D) for _ in range(NSETS)]
q = rnd(B, HQ, D)
att = torch.empty(B, HQ, D, device=dev, dtype=bf)
cu_q = torch.arange(B + 1, dtype=torch.int32, device=dev)
cu_k = (torch.arange(B + 1, dtype=torch.int32, device=dev) * cap).to(torch.int32)
used = torch.full((B,), CTX, dtype=torch.int32, device=dev)
tok = torch.empty(B, dtype=torch.int64, device=dev)
tok32 = torch.empty(B, dtype=torch.int32, device=dev)


def bench(name, fn, nbytes, iters=20):
    for i in range(4):
        fn(i % NSETS)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i in range(iters):
        ev[i][0].record()
        fn(i % NSETS)
        ev[i][1].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
    med = ts[len(ts) // 2]
    print(f"{name:34s} {med:8.1f} us  (min {ts[0]:7.1f})  {nbytes / med / 1e6:7.2f} TB/s of {nbytes / 1e6:8.1f} MB", flush=True)
    return med


tot = 0.0
tot += bench("qkv   [B,3584]x[4608,3584] bias", lambda i: ops.gemm(x, sets[i]["wqkv"], bias=sets[i]["bqkv"], out=qkv), 4608 * H * 2)
tot += bench(f"attn decode ctx={CTX} 28/4", lambda i: ops.attn_varlen(q, kbuf[i], vbuf[i], cu_q, cu_k, 1, cap, True, out=att, seqused_k=used),
             2 * B * CTX * HK * D * 2)
tot += bench("o     [B,3584]x[3584,3584] resid", lambda i: ops.gemm(x, sets[i]["wo"], resid=x, epilogue=ops.EPI_RESID, out=out), H * H * 2)
tot += bench("gate|up [B,3584]x[37888,3584] swiglu", lambda i: ops.gemm(x, sets[i]["wgu"], epilogue=ops.EPI_SWIGLU, out=act), 2 * I * H * 2)
tot += bench("down  [B,18944]x[3584,18944] resid", lambda i: ops.gemm(act, sets[i]["wd"], resid=x, epilogue=ops.EPI_RESID, out=out), H * I * 2)
print(f"layer GEMMs + attention: {tot:.1f} us -> x28 = {tot * 28 / 1e3:.2f} ms")
bench("lm_head [B,3584]x[152064,3584]", lambda i: ops.gemm(x, head, out=logits), V * H * 2, iters=8)
bench("argmax [B,152064]", lambda i: ops.argmax_rows(logits, tok, tok32), B * V * 2)
bench("rmsnorm [B,3584]", lambda i: ops.rmsnorm(x, sets[i]["bqkv"][:H], None, None, 1e-6, out=out), B * H * 4)
