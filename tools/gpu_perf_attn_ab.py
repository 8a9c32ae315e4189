"""A/B of the attention kernel's exp2 offload (BAGEL_ATTN_POLY, read once per process) on one box."""
import os, subprocess, sys
code = r'''
import sys, torch
sys.path.insert(0, ".")
from bagel_b200 import ops
from oracle import qwen2_mot as om
dev = "cuda"
def bench(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
torch.manual_seed(4)
# accuracy vs fp32 reference (small ragged case, causal and not)
for causal in (False, True):
    lq, lk = [100, 515, 1], [100, 700, 333]
    q = torch.randn(sum(lq), 28, 128, device=dev).to(torch.bfloat16); k = torch.randn(sum(lk), 4, 128, device=dev).to(torch.bfloat16); v = torch.randn(sum(lk), 4, 128, device=dev).to(torch.bfloat16)
    cq = torch.tensor([0] + torch.tensor(lq).cumsum(0).tolist(), dtype=torch.int32, device=dev); ck = torch.tensor([0] + torch.tensor(lk).cumsum(0).tolist(), dtype=torch.int32, device=dev)
    out = ops.attn_varlen(q, k, v, cq, ck, max(lq), max(lk), causal).float().cpu()
    ref = om.varlen_attention(q.cpu(), k.cpu(), v.cpu(), lq, lk, causal).float()
    print(f"  accuracy causal={causal}: max|err| {(out-ref).abs().max().item():.3e} mean {(out-ref).abs().mean().item():.3e}", flush=True)
lq = [4098] * 16; lk = [4164] * 16
q = torch.randn(sum(lq), 28, 128, device=dev).to(torch.bfloat16); k = torch.randn(sum(lk), 4, 128, device=dev).to(torch.bfloat16); v = torch.randn(sum(lk), 4, 128, device=dev).to(torch.bfloat16)
cq = torch.tensor([0] + torch.tensor(lq).cumsum(0).tolist(), dtype=torch.int32, device=dev); ck = torch.tensor([0] + torch.tensor(lk).cumsum(0).tolist(), dtype=torch.int32, device=dev)
out = torch.empty_like(q)
t = bench(lambda: ops.attn_varlen(q, k, v, cq, ck, 4098, 4164, False, out=out))
print(f"  denoise B=16 q=4098 kv=4164 28/4: {t:.3f} ms = {4.0*16*4098*4164*28*128/t/1e9:.0f} TFLOP/s", flush=True)
L = 16384
q = torch.randn(L, 32, 128, device=dev).to(torch.bfloat16); k = torch.randn(L, 32, 128, device=dev).to(torch.bfloat16); v = torch.randn(L, 32, 128, device=dev).to(torch.bfloat16)
cu = torch.tensor([0, L], dtype=torch.int32, device=dev); out = torch.empty_like(q)
for causal in (False, True):
    t = bench(lambda: ops.attn_varlen(q, k, v, cu, cu, L, L, causal, out=out))
    print(f"  L=16k MHA causal={causal}: {t:.3f} ms = {4.0*L*L*32*128/(2 if causal else 1)/t/1e9:.0f} TFLOP/s", flush=True)
'''
for v in ("0", "2", "3", "4"):
    print(f"BAGEL_ATTN_POLY={v}", flush=True)
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, BAGEL_ATTN_POLY=v))
