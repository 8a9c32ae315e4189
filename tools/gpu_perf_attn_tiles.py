"""One Q tile vs two Q tiles per CTA of attn_varlen_kernel<128> with the same K/V stream: tells how much of the
softmax time of one tile is hidden behind the other tile's MMAs (ideal: two tiles cost the same as one).
Lq=128 -> only tile 0 of each CTA is active; Lq=256 -> both. grid = (1, 28, B) in both cases."""
import sys
import torch
sys.path.insert(0, ".")
from bagel_b200 import ops

dev, bf = "cuda", torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
B, Hq, Hk, D, Lk = 32, 28, 4, 128, 8192
k = torch.randn(B * Lk, Hk, D, device=dev, generator=g).to(bf)
v = torch.randn(B * Lk, Hk, D, device=dev, generator=g).to(bf)
ck = (torch.arange(B + 1, device=dev) * Lk).to(torch.int32)
for Lq in (128, 256):
    q = torch.randn(B * Lq, Hq, D, device=dev, generator=g).to(bf)
    cq = (torch.arange(B + 1, device=dev) * Lq).to(torch.int32)
    out = torch.empty_like(q)
    for _ in range(3):
        ops.attn_varlen(q, k, v, cq, ck, Lq, Lk, False, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.attn_varlen(q, k, v, cq, ck, Lq, Lk, False, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    ctas = B * Hq
    waves = ctas / 148.0
    nblk = Lk // 128
    per_blk_us = t * 1e3 / (waves * nblk)
    fl = 4.0 * B * Lq * Lk * Hq * D
    print(f"Lq={Lq}: {t:.3f} ms, {fl / t / 1e9:.0f} TFLOP/s, {ctas} CTAs = {waves:.2f} waves x {nblk} KV blocks -> "
          f"{per_blk_us:.3f} us per CTA per KV block ({'1 tile' if Lq == 128 else '2 tiles'})", flush=True)
