"""Timeline of one attention CTA (clock64 stamps written by the kernel when BAGEL_ATTN_TRACE=<device ptr>)."""
import os, sys, torch
sys.path.insert(0, ".")
dev = "cuda"
trace = torch.zeros(3 * 64 * 8, dtype=torch.int64, device=dev)
os.environ["BAGEL_ATTN_TRACE"] = str(trace.data_ptr())
from bagel_b200 import ops
torch.manual_seed(0)
lq = [4098] * 16; lk = [4164] * 16
q = torch.randn(sum(lq), 28, 128, device=dev).to(torch.bfloat16); k = torch.randn(sum(lk), 4, 128, device=dev).to(torch.bfloat16); v = torch.randn(sum(lk), 4, 128, device=dev).to(torch.bfloat16)
cq = torch.tensor([0] + torch.tensor(lq).cumsum(0).tolist(), dtype=torch.int32, device=dev); ck = torch.tensor([0] + torch.tensor(lk).cumsum(0).tolist(), dtype=torch.int32, device=dev)
def bench(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
tt_ms = bench(lambda: ops.attn_varlen(q, k, v, cq, ck, 4098, 4164, False))
print(f"BAGEL_ATTN_POLY={os.environ.get('BAGEL_ATTN_POLY','default')}: {tt_ms:.3f} ms = {4.0*16*4098*4164*28*128/tt_ms/1e9:.0f} TFLOP/s (with trace stores on CTA 0)")
torch.cuda.synchronize()
t = trace.cpu().reshape(3, 64, 8)
t0 = int(t[t > 0].min())
nb = 33
print("softmax tile t, block j: wait_S | ld | max+alpha(+rescale) | exp+st issue | st_wait+arrive   (cycles); abs start")
for tt in range(2):
    for j in range(8, 14):
        r = t[tt, j]
        print(f"  t{tt} j{j:2d}: waitS {int(r[1]-r[0]):5d} ld {int(r[2]-r[1]):5d} max {int(r[3]-r[2]):5d} exp {int(r[4]-r[3]):5d} fin {int(r[5]-r[4]):5d} | start {int(r[0]-t0):7d} S_ready {int(r[1]-t0):7d} P_arrive {int(r[5]-t0):7d}")
print("MMA thread: waitP | issue PV | issue QK ; abs")
for j in range(8, 14):
    for tt in range(2):
        r = t[2, j, tt * 4: tt * 4 + 4]
        print(f"  j{j:2d} t{tt}: waitP {int(r[1]-r[0]):5d} issuePV {int(r[2]-r[1]):5d} issueQK {int(r[3]-r[2]):5d} | start {int(r[0]-t0):7d} P_seen {int(r[1]-t0):7d} done {int(r[3]-t0):7d}")
per = [int(t[0, j + 1, 0] - t[0, j, 0]) for j in range(4, 30)]
print("tile0 period per block (cycles):", per)
