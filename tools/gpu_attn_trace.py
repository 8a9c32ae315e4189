"""Timeline of the persistent attention kernel (clock64 stamps of CTA 0): how long the softmax of one 128 x 128 block takes, how
long the MMA lane waits for P, how long the softmax warps wait for S. Builds (here, with nvcc) a SEPARATE library with
-DBAGEL_ATTN_TRACE next to this file (tools/_trace/, git-ignored; it travels to the GPU box) and runs it there:

    python tools/gpu_attn_trace.py build        # in the container (no GPU needed)
    python tools/gpu_attn_trace.py run [L] [causal]     # on the GPU box
"""
import ctypes, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "tools" / "_trace" / "libbagel_b200_attntrace.so"


def build():
    sys.path.insert(0, str(ROOT))
    from bagel_b200 import build as bb
    bb.build()
    OUT.parent.mkdir(parents=True, exist_ok=True)
    obj = OUT.parent / "attn_trace.o"
    subprocess.check_call([bb._nvcc(), *bb.NVCC_FLAGS, "-DBAGEL_ATTN_TRACE", "-c", str(bb.CSRC / "attn.cu"), "-o", str(obj)])
    objs = [str(o) for o in sorted((bb.PKG_DIR / "build").glob("*.o")) if o.name != "attn.o"] + [str(obj)]
    subprocess.check_call([bb._nvcc(), "-shared", "-o", str(OUT), *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"])
    print("built", OUT)


def run(L=4096, causal=0, B=4, Hq=28, Hk=4, D=128):
    import numpy as np, torch
    lib = ctypes.CDLL(str(OUT))
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.randn(B * L, Hq, D, device=dev, dtype=torch.bfloat16, generator=g)
    k = torch.randn(B * L, Hk, D, device=dev, dtype=torch.bfloat16, generator=g)
    v = torch.randn(B * L, Hk, D, device=dev, dtype=torch.bfloat16, generator=g)
    o = torch.empty_like(q)
    cu = torch.arange(0, (B + 1) * L, L, dtype=torch.int32, device=dev)
    P = ctypes.c_void_p
    fn = lib.bagel_attn_varlen_fwd
    fn.restype = ctypes.c_int
    fn.argtypes = [P, P, P, P, P, P] + [ctypes.c_int] * 9 + [ctypes.c_float] + [ctypes.c_longlong] * 4 + [P, P]
    def call():
        rc = fn(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), cu.data_ptr(), cu.data_ptr(), B * L, B * L, B, Hq, Hk, D, L, L,
                causal, D ** -0.5, Hq * D, Hk * D, Hk * D, Hq * D, None, None)
        assert rc == 0, rc
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); call(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    fl = 4.0 * B * L * L * Hq * D * (0.5 if causal else 1.0)
    print(f"L={L} causal={causal}: {ms:.3f} ms = {fl / ms / 1e9:.0f} TFLOP/s (traced build)")
    n = 3 * 1024 * 8
    buf = (ctypes.c_longlong * n)()
    lib.bagel_attn_trace_read.argtypes = [P, ctypes.c_int]
    assert lib.bagel_attn_trace_read(buf, n) == 0
    t = np.frombuffer(buf, dtype=np.int64).reshape(3, 1024, 8)
    lo = 40               # steady-state iterations: skip the ramp, stop where CTA 0 ran out of work (unwritten entries are 0)
    for tile in (0, 1):
        hi = max(lo + 2, int((t[tile, :, 1] > 0).sum()) - 4)
        e = t[tile, lo:hi]
        per = np.diff(e[:, 1])
        print(f"softmax tile {tile}: period {np.median(per):.0f} clk | wait for S {np.median(e[:,1]-e[:,0]):.0f} | "
              f"S ready -> exps issued {np.median(e[:,2]-e[:,1]):.0f} | -> P stored {np.median(e[:,3]-e[:,2]):.0f} | "
              f"-> arrive {np.median(e[:,4]-e[:,3]):.0f} | arrive -> next wait {np.median(e[1:,0]-e[:-1,4]):.0f}")
    hi = max(lo + 2, int((t[2, :, 1] > 0).sum()) // 2 - 4)
    m = t[2, 2 * lo:2 * hi]
    for tile in (0, 1):
        mm = m[m[:, 3] == tile]
        print(f"MMA lane, tile {tile}: wait for P {np.median(mm[:,1]-mm[:,0]):.0f} clk | issue PV+QK {np.median(mm[:,2]-mm[:,1]):.0f} | "
              f"period {np.median(np.diff(mm[:,1])):.0f}")
    m0_, m1_ = m[m[:, 3] == 0], m[m[:, 3] == 1]
    k = min(len(m0_), len(m1_)) - 1
    if k > 10:
        print(f"MMA lane per key block: wait for V_j + K_j+1 {np.median(m0_[:k,5]-m0_[:k,4]):.0f} clk | commits after the 2nd tile's issue "
              f"{np.median(m1_[:k,6]-m1_[:k,2]):.0f} | loop turn-around {np.median(m0_[1:k+1,4]-m1_[:k,6]):.0f}")
    # latency from the softmax arrive to the MMA lane seeing it, and from MMA issue to the softmax seeing S
    s0 = t[0, lo:max(lo + 2, int((t[0, :, 1] > 0).sum()) - 4)]
    m0 = m[m[:, 3] == 0]
    # align by nearest following stamp
    arr = s0[:, 4]
    seen = m0[:, 1]
    d = [seen[seen >= x][0] - x for x in arr[:200] if (seen >= x).any()]
    print(f"softmax arrive -> MMA lane past the wait (tile 0): median {np.median(d):.0f} clk")
    iss = m0[:, 2]
    sr = s0[:, 1]
    d2 = [sr[sr >= x][0] - x for x in iss[:200] if (sr >= x).any()]
    print(f"MMA issue done -> softmax sees S (tile 0): median {np.median(d2):.0f} clk (includes PV + QK execution, 1024 clk ideal)")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        args = [int(x) for x in sys.argv[2:]]
        run(*args)
