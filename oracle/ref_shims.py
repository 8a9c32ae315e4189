"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference (ByteDance-Seed/Bagel, mounted read-only at
/root/reference) on CPU, with the three harness-side shims SURVEY.md §8(c) lists. Used only by
tests/golden/make_golden.py (to generate the committed fixtures) and by the optional live cross-check tests
that skip when /root/reference is absent (it does not exist on the GPU box).

Shims (the reference itself is untouched):
  1. transformers>=5 dropped ROPE_INIT_FUNCTIONS["default"], which modeling/qwen2/modeling_qwen2.py:105 indexes;
  2. Qwen2Config no longer defaults pad_token_id (read at modeling/bagel/qwen2_navit.py:946);
  3. flash_attn_varlen_func has no CPU kernel -> per-sequence fp32 SDPA on the bf16 inputs, bottom-right
     aligned causal mask, GQA by head repetition (the documented semantics of flash-attn >= 2.1).
  4. (training forward only) `sdpa_kernel(backends=[EFFICIENT_ATTENTION])` around the masked SDPA call of
     PackedAttentionMoT.forward_train (qwen2_navit.py:462) selects a CUDA kernel and leaves no viable backend on the
     CPU -> replaced by a no-op context manager, i.e. torch's default CPU SDPA on the same operands.
"""
from __future__ import annotations

import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("BAGEL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "modeling", "bagel"))


def cpu_varlen_attention(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q=None, max_seqlen_k=None,
                         causal=False, **_unused):
    """q [Sq,Hq,D], k/v [Sk,Hk,D] bf16 -> [Sq,Hq,D]; softmax in fp32, output cast back to q.dtype.
    Runs with autocast disabled: the reference calls this under torch.autocast, which would silently turn
    the fp32 matmuls below into bf16 ones (flash-attn accumulates in fp32 regardless of autocast)."""
    with torch.autocast("cpu", enabled=False):
        return _cpu_varlen_attention(q, k, v, cu_seqlens_q, cu_seqlens_k, causal)


def _cpu_varlen_attention(q, k, v, cu_seqlens_q, cu_seqlens_k, causal):
    out = torch.empty_like(q)
    hq, hk = q.shape[1], k.shape[1]
    rep = hq // hk
    scale = q.shape[-1] ** -0.5
    nb = cu_seqlens_q.numel() - 1
    for b in range(nb):
        qs, qe = int(cu_seqlens_q[b]), int(cu_seqlens_q[b + 1])
        ks, ke = int(cu_seqlens_k[b]), int(cu_seqlens_k[b + 1])
        if qe == qs:
            continue
        qb = q[qs:qe].float().transpose(0, 1)                      # [Hq,Lq,D]
        kb = k[ks:ke].float().transpose(0, 1).repeat_interleave(rep, dim=0)
        vb = v[ks:ke].float().transpose(0, 1).repeat_interleave(rep, dim=0)
        s = torch.matmul(qb, kb.transpose(1, 2)) * scale           # [Hq,Lq,Lk]
        if causal:
            lq, lk = qe - qs, ke - ks
            mask = torch.ones(lq, lk, dtype=torch.bool).tril(diagonal=lk - lq)
            s = s.masked_fill(~mask, float("-inf"))
        p = torch.softmax(s, dim=-1)
        out[qs:qe] = torch.matmul(p, vb).transpose(0, 1).to(q.dtype)
    return out


_loaded = None


def load_reference():
    """Returns a namespace of the reference modules with shims applied. Raises if the tree is absent."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    from transformers import modeling_rope_utils as mru

    if "default" not in mru.ROPE_INIT_FUNCTIONS:
        def _default_rope(config, device=None, seq_len=None, **kw):
            base = config.rope_theta
            dim = getattr(config, "head_dim", None) or config.hidden_size // config.num_attention_heads
            inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).float().to(device) / dim))
            return inv_freq, 1.0
        mru.ROPE_INIT_FUNCTIONS["default"] = _default_rope

    import types

    ns = types.SimpleNamespace()
    import modeling.bagel.qwen2_navit as qn
    import modeling.bagel.siglip_navit as sn
    import modeling.bagel.bagel as bg
    import modeling.bagel.modeling_utils as mu
    import modeling.autoencoder as ae
    import modeling.qwen2.modeling_qwen2 as mq
    import data.data_utils as du

    qn.flash_attn_varlen_func = cpu_varlen_attention
    sn.flash_attn_varlen_func = cpu_varlen_attention
    import contextlib
    qn.sdpa_kernel = lambda *a, **k: contextlib.nullcontext()
    ns.qwen2_navit, ns.siglip_navit, ns.bagel, ns.modeling_utils = qn, sn, bg, mu
    ns.autoencoder, ns.modeling_qwen2, ns.data_utils = ae, mq, du
    try:
        import inferencer as inf
        ns.inferencer = inf
    except Exception as e:  # PIL etc. present; keep optional
        ns.inferencer = None
        ns.inferencer_error = e
    _loaded = ns
    return ns


def make_llm_config(ns, **kw):
    """Qwen2Config with the post-load overrides every shipped loader applies (app.py:40-46)."""
    kw.setdefault("pad_token_id", None)
    kw.setdefault("qk_norm", True)
    kw.setdefault("tie_word_embeddings", False)
    kw.setdefault("layer_module", "Qwen2MoTDecoderLayer")
    cfg = ns.qwen2_navit.Qwen2Config(**kw)
    if getattr(cfg, "rope_theta", None) is None:
        cfg.rope_theta = kw.get("rope_theta", 1000000.0)
    return cfg


def cast_parameters(module, dtype):
    """bf16 weights the way app.py:111 gets them (accelerate `dtype=` casts checkpoint tensors only):
    parameters are cast, non-persistent buffers such as rotary inv_freq stay fp32."""
    for p in module.parameters():
        p.data = p.data.to(dtype)
    return module
