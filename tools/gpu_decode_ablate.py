"""In-graph cost of every kernel of the text-decode step (BAGEL-7B dims, batch 32, 1245-token context).

ncu launch lists time each kernel alone, cold and serialised (short kernels are dominated by launch / clock ramp there), and
CUDA events around single launches measure the host. This tool measures what a kernel costs INSIDE the replayed CUDA graph:
  (1) ablation: the 28-layer step graph with one kernel family removed (numerics are garbage, timing is not), and
  (2) a graph of 200 back-to-back launches of one small kernel (the floor a graph node costs).
Usage: python tools/gpu_decode_ablate.py [layers=28] [B=32] [ctx=1245] [quick]"""
import sys
import torch
sys.path.insert(0, ".")
from bagel_b200 import ops, synthetic

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 28
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 1245
quick = len(sys.argv) > 4 and sys.argv[4] == "quick"
dev, BF16 = "cuda", torch.bfloat16
model = synthetic.build_random_bagel(device=dev, seed=0, num_layers=layers)
lm = model.language_model.model
cfg = lm.config
L, H, Hq, Hk, D = cfg.num_hidden_layers, cfg.hidden_size, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
w = Hk * D
cap = ctx + 160
total = B * cap
kbuf = torch.zeros((L, total, w), dtype=BF16, device=dev)
vbuf = torch.zeros((L, total, w), dtype=BF16, device=dev)
kbuf.normal_(0, 1.0)
vbuf.normal_(0, 1.0)
k_begin = (torch.arange(B + 1, dtype=torch.int64) * cap).to(dev, torch.int32)
cu_q = torch.arange(B + 1, dtype=torch.int32, device=dev)
seq_len = torch.full((B,), ctx, dtype=torch.int32, device=dev)
pos = torch.full((B,), ctx, dtype=torch.int64, device=dev)
tokens32 = torch.randint(0, 1000, (B,), dtype=torch.int32, device=dev)
kv_rows = torch.zeros(B, dtype=torch.int32, device=dev)
seqused = torch.zeros(B, dtype=torch.int32, device=dev)
x = torch.empty((B, H), dtype=BF16, device=dev)
logits = torch.empty((B, cfg.vocab_size), dtype=BF16, device=dev)
eps = cfg.rms_norm_eps
bufs = dict(xb=torch.empty_like(x), h=torch.randn((B, H), device=dev).to(BF16),
            qkv=torch.randn((B, (Hq + 2 * Hk) * D), device=dev).to(BF16),
            q=torch.randn((B, Hq * D), device=dev).to(BF16), att=torch.randn((B, Hq * D), device=dev).to(BF16),
            act=torch.randn((B, cfg.intermediate_size), device=dev).to(BF16), out=torch.empty_like(x))
cos = torch.empty((B, D // 2), dtype=torch.float32, device=dev)
sin = torch.empty((B, D // 2), dtype=torch.float32, device=dev)
head = model.language_model.lm_head


def body(skip=()):
    ops.copy_rows(lm.embed_tokens.weight, x, src_rows=tokens32, M=B)
    ops.rope_table_into(pos, lm.inv_freq, cos, sin, True)
    ops.decode_prepare(k_begin, seq_len, kv_rows, seqused)
    xa, xb, h = x, bufs["xb"], bufs["h"]
    for li, layer in enumerate(lm.layers):
        e = layer.und
        if "norm" not in skip:
            ops.rmsnorm(xa, e.ln_in, None, None, eps, out=h)
        if "qkv" not in skip:
            ops.gemm(h, e.wqkv, bias=e.bqkv, out=bufs["qkv"])
        if "qkrope" not in skip:
            ops.qk_norm_rope(bufs["qkv"], e.q_norm, e.k_norm, None, None, None, cos, sin, bufs["q"], kbuf[li], vbuf[li], kv_rows,
                             Hq, Hk, D, eps, False)
        if "attn" not in skip:
            ops.attn_varlen(bufs["q"].view(B, Hq, D), kbuf[li].view(-1, Hk, D), vbuf[li].view(-1, Hk, D), cu_q, k_begin, 1, cap,
                            True, out=bufs["att"].view(B, Hq, D), seqused_k=seqused)
        if "o" not in skip:
            ops.gemm(bufs["att"], e.wo, resid=xa, epilogue=ops.EPI_RESID, out=xb)
        if "norm" not in skip:
            ops.rmsnorm(xb, e.ln_post, None, None, eps, out=h)
        if "gu" not in skip:
            ops.gemm(h, e.wgu, epilogue=ops.EPI_SWIGLU, out=bufs["act"])
        if "down" not in skip:
            ops.gemm(bufs["act"], e.wd, resid=xb, epilogue=ops.EPI_RESID, out=xa)
    if "head" not in skip:
        ops.rmsnorm(xa, lm.norm, None, None, eps, out=bufs["out"])
        ops.gemm(bufs["out"], head.weight, bias=head.bias, out=logits)


def time_graph(fn, reps=15):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


full = time_graph(lambda: body())
print(f"full step ({L} layers, B={B}, ctx={ctx}): {full:.3f} ms", flush=True)
for name in (("norm", "qkrope", "attn") if quick else ("norm", "qkv", "qkrope", "attn", "o", "gu", "down", "head")):
    t = time_graph(lambda: body((name,)))
    per = (full - t) * 1e3 / (L * (2 if name == "norm" else 1)) if name != "head" else (full - t) * 1e3
    print(f"  without {name:7s}: {t:.3f} ms  -> marginal {full - t:.3f} ms  ({per:.1f} us per launch)", flush=True)
for combo in (() if quick else (("norm", "qkrope"), ("qkv", "o", "gu", "down", "head"), ("norm", "qkrope", "attn"))):
    t = time_graph(lambda: body(combo))
    print(f"  without {'+'.join(combo)}: {t:.3f} ms -> marginal {full - t:.3f} ms", flush=True)

# floor of a graph node: 200 back-to-back launches of one small kernel
e = lm.layers[0].und
h = bufs["h"]
small = {
    "rmsnorm [B,3584]": lambda: ops.rmsnorm(x, e.ln_in, None, None, eps, out=h),
    "qk_norm_rope": lambda: ops.qk_norm_rope(bufs["qkv"], e.q_norm, e.k_norm, None, None, None, cos, sin, bufs["q"], kbuf[0], vbuf[0],
                                             kv_rows, Hq, Hk, D, eps, False),
    "attn decode": lambda: ops.attn_varlen(bufs["q"].view(B, Hq, D), kbuf[0].view(-1, Hk, D), vbuf[0].view(-1, Hk, D), cu_q, k_begin, 1,
                                           cap, True, out=bufs["att"].view(B, Hq, D), seqused_k=seqused),
    "decode_prepare": lambda: ops.decode_prepare(k_begin, seq_len, kv_rows, seqused),
}
for name, fn in small.items():
    def many():
        for _ in range(200):
            fn()
    t = time_graph(many, reps=7)
    print(f"  200 x {name}: {t * 1e3 / 200:.2f} us per launch in a graph", flush=True)
