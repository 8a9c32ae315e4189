"""GPU diagnostic for bagel_attn_varlen_fwd and the elementwise kernels (vs torch fp32 / oracle)."""
import sys, math
import torch
sys.path.insert(0, ".")
from bagel_b200 import ops
from oracle import qwen2_mot as om

torch.manual_seed(0)
dev = "cuda"

def ref_attn(q, k, v, lq, lk, causal):
    out = torch.empty_like(q, dtype=torch.float32)
    rep = q.shape[1] // k.shape[1]; scale = q.shape[-1] ** -0.5
    qs = ks = 0
    for a, b in zip(lq, lk):
        if a:
            qb = q[qs:qs+a].float().transpose(0, 1); kb = k[ks:ks+b].float().transpose(0, 1).repeat_interleave(rep, 0)
            vb = v[ks:ks+b].float().transpose(0, 1).repeat_interleave(rep, 0)
            s = qb @ kb.transpose(1, 2) * scale
            if causal:
                keep = torch.ones(a, b, dtype=torch.bool, device=q.device).tril(diagonal=b - a)
                s = s.masked_fill(~keep, float("-inf"))
            out[qs:qs+a] = (torch.softmax(s, -1) @ vb).transpose(0, 1)
        qs += a; ks += b
    return out

def report(name, out, ref, atol, rtol):
    out = out.float(); ref = ref.float()
    err = (out - ref).abs(); tol = atol + rtol * ref.abs()
    bad = int((err > tol).sum())
    print(f"[{name}] max_abs_err={err.max().item():.4e} mean_abs_err={err.mean().item():.3e} bad={bad}/{out.numel()} finite={bool(torch.isfinite(out).all())}", flush=True)
    return bad == 0

ok = True
cases = [  # (lq list, lk list, Hq, Hk, D, causal)
    ([128], [128], 1, 1, 128, False),
    ([256], [256], 2, 1, 128, False),
    ([300], [300], 4, 2, 128, False),
    ([512], [512], 4, 2, 64, True),
    ([130], [642], 4, 2, 64, False),
    ([100, 515, 1], [100, 700, 333], 28, 4, 128, False),
    ([100, 515, 1], [100, 700, 333], 28, 4, 128, True),
    ([4098, 4098], [4164, 4098], 28, 4, 128, False),
    ([729, 729, 300], [729, 729, 300], 16, 16, 64, False),
]
for (lq, lk, Hq, Hk, D, causal) in cases:
    Sq, Sk = sum(lq), sum(lk)
    q = torch.randn(Sq, Hq, D, device=dev).to(torch.bfloat16)
    k = torch.randn(Sk, Hk, D, device=dev).to(torch.bfloat16)
    v = torch.randn(Sk, Hk, D, device=dev).to(torch.bfloat16)
    cq = torch.tensor([0] + list(torch.tensor(lq).cumsum(0)), dtype=torch.int32, device=dev)
    ck = torch.tensor([0] + list(torch.tensor(lk).cumsum(0)), dtype=torch.int32, device=dev)
    try:
        out = ops.attn_varlen(q, k, v, cq, ck, max(lq), max(lk), causal)
        torch.cuda.synchronize()
    except Exception as e:
        print(f"[attn lq={lq} D={D}] EXCEPTION {e}", flush=True); ok = False; break
    good = report(f"attn lq={lq} lk={lk} H={Hq}/{Hk} D={D} causal={causal}", out, ref_attn(q, k, v, lq, lk, causal), 2e-2, 2e-2)
    if not good:
        o = out.float(); r = ref_attn(q, k, v, lq, lk, causal)
        e = (o - r).abs()
        rows_bad = (e.amax(dim=(1, 2)) > 5e-2).nonzero().flatten()[:10].tolist()
        print("   bad rows (first 10):", rows_bad, " per-head max err:", [round(x, 3) for x in e.amax(dim=(0, 2)).tolist()[:8]])
        print("   col-block max err:", [round(e[..., c*16:(c+1)*16].max().item(), 3) for c in range(D // 16)])
        print("   sample out/ref:", o[0, 0, :4].tolist(), r[0, 0, :4].tolist())
    ok &= good

# ---- elementwise kernels vs oracle (CPU) ----
N, H = 515, 3584
x = torch.randn(N, H, device=dev).to(torch.bfloat16)
w0 = (1 + 0.1 * torch.randn(H, device=dev)).to(torch.bfloat16); w1 = (1 + 0.1 * torch.randn(H, device=dev)).to(torch.bfloat16)
ex = (torch.rand(N, device=dev) > 0.3).to(torch.uint8)
y = ops.rmsnorm(x, w0, w1, ex); torch.cuda.synchronize()
xc, exc = x.cpu(), ex.cpu().bool()
ref = torch.where(exc[:, None], om.rms_norm(xc, w1.cpu(), 1e-6), om.rms_norm(xc, w0.cpu(), 1e-6))
ok &= report("rmsnorm routed", y.cpu(), ref, 1e-6, 8e-3)

for D, Hq, Hk in ((128, 28, 4), (64, 4, 2)):
    for flow in (1, 0):
        N = 300
        qkv = torch.randn(N, (Hq + 2 * Hk) * D, device=dev).to(torch.bfloat16)
        qw = [(1 + 0.1 * torch.randn(D, device=dev)).to(torch.bfloat16) for _ in range(2)]
        kw = [(1 + 0.1 * torch.randn(D, device=dev)).to(torch.bfloat16) for _ in range(2)]
        ex = (torch.rand(N, device=dev) > 0.3).to(torch.uint8)
        pos = torch.randint(0, 5000, (N,), device=dev, dtype=torch.int64)
        inv_freq = (1.0 / (1e6 ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))).to(dev)
        cos, sin = ops.rope_table(pos, inv_freq, True)
        c_ref, s_ref = om.rope_tables(pos.cpu(), D, 1e6, torch.bfloat16)
        ok &= report(f"rope_table D={D}", cos.cpu(), c_ref[:, :D // 2], 1e-6, 8e-3)
        q_out = torch.zeros(N, Hq * D, device=dev, dtype=torch.bfloat16)
        kbuf = torch.zeros(N + 50, Hk * D, device=dev, dtype=torch.bfloat16); vbuf = torch.zeros_like(kbuf)
        rows = (torch.randperm(N + 50, device=dev)[:N]).to(torch.int32)
        ops.qk_norm_rope(qkv, qw[0], kw[0], qw[1], kw[1], ex, cos, sin, q_out, kbuf, vbuf, rows, Hq, Hk, D, 1e-6, bool(flow))
        torch.cuda.synchronize()
        # oracle restatement of qwen2_navit.py:542-557 / 518-519 on CPU
        qc = qkv.cpu(); exc = ex.cpu().bool()
        q = qc[:, :Hq * D].reshape(N, Hq, D); k = qc[:, Hq * D:(Hq + Hk) * D].reshape(N, Hk, D); v = qc[:, (Hq + Hk) * D:].reshape(N, Hk, D)
        if flow:
            q = q.float(); k = k.float()
        def nrm(t, w_und, w_gen):
            return torch.where(exc[:, None, None], om.rms_norm(t, w_gen.cpu(), 1e-6), om.rms_norm(t, w_und.cpu(), 1e-6))
        qn = nrm(q, qw[0], qw[1]); kn = nrm(k, kw[0], kw[1])
        qr, kr = om.apply_rope(qn, kn, c_ref, s_ref)
        ok &= report(f"qk_norm_rope q D={D} flow={flow}", q_out.cpu().reshape(N, Hq, D), qr.to(torch.bfloat16), 1e-6, 8e-3)
        ok &= report(f"qk_norm_rope k D={D} flow={flow}", kbuf[rows.long()].cpu().reshape(N, Hk, D), kr.to(torch.bfloat16), 1e-6, 8e-3)
        ok &= report(f"qk_norm_rope v D={D} flow={flow}", vbuf[rows.long()].cpu().reshape(N, Hk, D), v, 0, 0)

# cfg + euler vs a torch restatement of bagel.py:873-907,746
M, Cc = 1000, 64
for rt in ("global", "channel", "text_channel"):
    for sI in (1.0, 1.5):
        v = torch.randn(M + 20, Cc, device=dev).to(torch.bfloat16); vT = torch.randn(M + 20, Cc, device=dev).to(torch.bfloat16); vI = torch.randn(M + 20, Cc, device=dev).to(torch.bfloat16)
        rows = torch.arange(10, 10 + M, device=dev, dtype=torch.int32)
        x = torch.randn(M, Cc, device=dev); x0 = x.clone()
        ws = torch.zeros(2, device=dev)
        ops.cfg_euler_step(v, vT, vI if sI > 1 else None, rows, x, ws, 4.0, sI, 0.0, rt, 0.037)
        torch.cuda.synchronize()
        v_, vT_, vI_ = v[10:10 + M].cpu(), vT[10:10 + M].cpu(), vI[10:10 + M].cpu()
        sT = 4.0
        if rt == "text_channel":
            u = vT_ + sT * (v_ - vT_)
            sc = (torch.norm(v_, dim=-1, keepdim=True) / (torch.norm(u, dim=-1, keepdim=True) + 1e-8)).clamp(min=0.0, max=1.0)
            ut = u * sc
            w = vI_ + sI * (ut - vI_) if sI > 1 else ut
        else:
            u = vT_ + sT * (v_ - vT_)
            w_ = vI_ + sI * (u - vI_) if sI > 1 else u
            if rt == "global":
                nv, nw = torch.norm(v_), torch.norm(w_)
            else:
                nv, nw = torch.norm(v_, dim=-1, keepdim=True), torch.norm(w_, dim=-1, keepdim=True)
            sc = (nv / (nw + 1e-8)).clamp(min=0.0, max=1.0)
            w = w_ * sc
        xr = x0.cpu() - w * torch.tensor(0.037)
        ok &= report(f"cfg_euler {rt} sI={sI}", x.cpu(), xr, 2e-3, 1e-3)

print("ALL_OK" if ok else "SOME_FAILED")
