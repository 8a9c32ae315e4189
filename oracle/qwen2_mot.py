"""TEST INFRASTRUCTURE — CPU restatement of the packed MoT Qwen2 language model forward
(reference: modeling/bagel/qwen2_navit.py:499-600, 757-831, 1018-1092; modeling/qwen2/modeling_qwen2.py:45-201).

Pure functions over `sd`, a dict with the reference's parameter names (prefix "model." for the decoder,
"lm_head.weight"). Dtype behaviour follows the reference under `torch.autocast(bfloat16)`: every nn.Linear
runs with bf16 operands and returns bf16; elementwise ops follow torch type promotion. `sd` may hold bf16
weights ("mode A", app.py:111) or fp32 weights ("mode B", eval/gen/gen_images_mp.py:159-175).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

BF16 = torch.bfloat16
_AUTOCAST = [torch.bfloat16]  # dtype nn.Linear / attention run in; see high_precision()
# Optional replacement for the attention contraction: callable (q, k, v, q_lens, k_lens, causal) -> out, or None for
# the restatement below (which is what is pinned bit-for-bit against the reference's CPU run). oracle/gpu_leg.py
# installs the real flash_attn_varlen_func here to form the "reference PyTorch path on the GPU" leg.
_ATTN_IMPL = [None]


class attention_impl:
    """Context manager: run the oracle with another varlen attention (e.g. flash-attn on CUDA)."""

    def __init__(self, fn):
        self.fn = fn

    def __enter__(self):
        self.prev = _ATTN_IMPL[0]
        _ATTN_IMPL[0] = self.fn

    def __exit__(self, *a):
        _ATTN_IMPL[0] = self.prev


class high_precision:
    """Context manager: run the same graph in fp32 end to end (fp32 weights expected) — the "exact" answer the
    bf16 pipelines (reference and GPU build) both approximate. Used by the parity tests to express tolerances
    as "no further from the truth than the reference itself"."""

    def __enter__(self):
        _AUTOCAST[0] = torch.float32

    def __exit__(self, *a):
        _AUTOCAST[0] = torch.bfloat16


class LazyF32(dict):
    """State dict whose values are upcast to fp32 on access: the fp32-truth evaluation of a 7B model without holding a
    second, 57 GB fp32 copy of the weights (use together with high_precision())."""

    def __getitem__(self, k):
        return dict.__getitem__(self, k).float()


@dataclass
class LMConfig:
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    vocab_size: int
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1000000.0
    # decoder layer class (qwen2_navit.py:936-940): "Qwen2MoTDecoderLayer" (every module duplicated for the gen expert),
    # "Qwen2MoEDecoderLayer" (:834-933 — shared attention / norms, only the MLP duplicated), "Qwen2DecoderLayer" (dense)
    layer_module: str = "Qwen2MoTDecoderLayer"

    @property
    def use_moe(self) -> bool:
        return "Mo" in self.layer_module          # qwen2_navit.py:948

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


def linear(x, w, b=None):
    """nn.Linear under autocast(bf16): operands cast to bf16, bf16 result."""
    dt = _AUTOCAST[0]
    return F.linear(x.to(dt), w.to(dt), None if b is None else b.to(dt))


def rms_norm(x, w, eps):
    """modeling_qwen2.py:54-59 — cast back to the input dtype BEFORE the weight multiply."""
    dt = x.dtype
    x32 = x.to(torch.float32)
    var = x32.pow(2).mean(-1, keepdim=True)
    x32 = x32 * torch.rsqrt(var + eps)
    return w * x32.to(dt)


def rope_tables(position_ids, head_dim, theta, dtype):
    """modeling_qwen2.py:130-150 — fp32 angles, halves duplicated, then cast to the hidden-stream dtype."""
    inv_freq = (1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))).to(position_ids.device)
    freqs = (inv_freq[None, :, None].float() @ position_ids[None, None, :].float()).transpose(1, 2)[0]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def apply_rope(q, k, cos, sin):
    """modeling_qwen2.py:162-186 with unsqueeze_dim=1 (heads)."""
    c, s = cos.unsqueeze(1), sin.unsqueeze(1)
    return q * c + _rot_half(q) * s, k * c + _rot_half(k) * s


def swiglu_mlp(x, sd, pfx):
    """modeling_qwen2.py:200-201."""
    g = linear(x, sd[pfx + "gate_proj.weight"])
    u = linear(x, sd[pfx + "up_proj.weight"])
    return linear(F.silu(g) * u, sd[pfx + "down_proj.weight"])


def varlen_attention(q, k, v, q_lens, k_lens, causal):
    """Semantics of flash_attn_varlen_func as the reference calls it (qwen2_navit.py:579-588): per-sample
    softmax(q k^T / sqrt(d)) v, GQA, bottom-right aligned causal mask, fp32 math on bf16 inputs, bf16 out."""
    if _ATTN_IMPL[0] is not None and _AUTOCAST[0] == torch.bfloat16:
        return _ATTN_IMPL[0](q, k, v, q_lens, k_lens, causal)
    out = torch.empty_like(q)
    rep = q.shape[1] // k.shape[1]
    scale = q.shape[-1] ** -0.5
    qs = ks = 0
    for lq, lk in zip(q_lens, k_lens):
        lq, lk = int(lq), int(lk)
        if lq:
            qb = q[qs:qs + lq].float().transpose(0, 1)
            kb = k[ks:ks + lk].float().transpose(0, 1).repeat_interleave(rep, dim=0)
            vb = v[ks:ks + lk].float().transpose(0, 1).repeat_interleave(rep, dim=0)
            s = torch.matmul(qb, kb.transpose(1, 2)) * scale
            if causal:
                keep = torch.ones(lq, lk, dtype=torch.bool, device=q.device).tril(diagonal=lk - lq)
                s = s.masked_fill(~keep, float("-inf"))
            out[qs:qs + lq] = torch.matmul(torch.softmax(s, dim=-1), vb).transpose(0, 1).to(q.dtype)
        qs += lq
        ks += lk
    return out


class KVCache:
    """Mirror of NaiveCache (qwen2_navit.py:207-221): per-layer packed [sum kv, Hk, d] tensors or None."""

    def __init__(self, num_layers):
        self.key_cache = {i: None for i in range(num_layers)}
        self.value_cache = {i: None for i in range(num_layers)}


def _attention(x, sd, cfg: LMConfig, li: int, cos, sin, query_lens, packed_query_indexes, cache: Optional[KVCache],
               key_values_lens, packed_key_value_indexes, update, is_causal, mode, vae_idx, text_idx):
    """qwen2_navit.py:499-600 (PackedAttentionMoT.forward_inference)."""
    p = f"model.layers.{li}.self_attn."
    Hq, Hk, d = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    eps = cfg.rms_norm_eps
    if cfg.layer_module != "Qwen2MoTDecoderLayer":
        mode = "und"        # dense PackedAttention (:313-378) is the und branch of PackedAttentionMoT, whatever the mode
    if mode == "und":
        q = linear(x, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"]).view(-1, Hq, d)
        k = linear(x, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"]).view(-1, Hk, d)
        v = linear(x, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"]).view(-1, Hk, d)
        q = rms_norm(q, sd[p + "q_norm.weight"], eps)
        k = rms_norm(k, sd[p + "k_norm.weight"], eps)
    else:
        x = x.to(_AUTOCAST[0])
        n = x.shape[0]
        q = x.new_zeros((n, Hq * d))
        k = x.new_zeros((n, Hk * d))
        v = x.new_zeros((n, Hk * d))
        xt, xv = x[text_idx], x[vae_idx]
        for dst, name in ((q, "q"), (k, "k"), (v, "v")):
            dst[text_idx] = linear(xt, sd[p + f"{name}_proj.weight"], sd[p + f"{name}_proj.bias"])
            dst[vae_idx] = linear(xv, sd[p + f"{name}_proj_moe_gen.weight"], sd[p + f"{name}_proj_moe_gen.bias"])
        q, k, v = q.view(-1, Hq, d), k.view(-1, Hk, d), v.view(-1, Hk, d)
        q = q.to(torch.float32)
        q[text_idx] = rms_norm(q[text_idx], sd[p + "q_norm.weight"], eps)
        q[vae_idx] = rms_norm(q[vae_idx], sd[p + "q_norm_moe_gen.weight"], eps)
        k = k.to(torch.float32)
        k[text_idx] = rms_norm(k[text_idx], sd[p + "k_norm.weight"], eps)
        k[vae_idx] = rms_norm(k[vae_idx], sd[p + "k_norm_moe_gen.weight"], eps)

    q, k = apply_rope(q, k, cos, sin)
    q, k, v = q.to(_AUTOCAST[0]), k.to(_AUTOCAST[0]), v.to(_AUTOCAST[0])

    if cache is not None and cache.key_cache[li] is not None:
        pk, pv = cache.key_cache[li], cache.value_cache[li]
        total = int(sum(query_lens)) + int(sum(key_values_lens))
        mk = pk.new_zeros((total, Hk, d))
        mv = pk.new_zeros((total, Hk, d))
        mk[packed_query_indexes] = k
        mk[packed_key_value_indexes] = pk
        mv[packed_query_indexes] = v
        mv[packed_key_value_indexes] = pv
        kv_lens = key_values_lens + query_lens
    else:
        mk, mv, kv_lens = k, v, query_lens

    o = varlen_attention(q, mk, mv, query_lens.tolist(), kv_lens.tolist(), is_causal)
    o = o.reshape(-1, Hq * d)
    if mode == "und":
        o = linear(o, sd[p + "o_proj.weight"])
    else:
        o[text_idx] = linear(o[text_idx], sd[p + "o_proj.weight"])
        o[vae_idx] = linear(o[vae_idx], sd[p + "o_proj_moe_gen.weight"])
    if update:
        cache.key_cache[li], cache.value_cache[li] = mk, mv
    return o


class TaylorSeerState:
    """Step-cache state of ONE velocity branch (main / text-CFG / image-CFG): restatement of
    modeling/cache_utils/taylorseer.py — `cache_init` (:128-166: fresh_threshold 3, max_order 6, first_enhance 5,
    taylor_cache True, fresh_ratio 0), `cal_type` (:80-122), `force_scheduler` (:62-76, linear_step_weight 0 =>
    cal_threshold = round(3 / 1) = 3), `derivative_approximation` (:12-32), `taylor_formula` (:34-47) — as driven by
    Qwen2Model.forward_inference (qwen2_navit.py:1034-1037, 1057-1061, 1086-1087) and the decoder layer
    (qwen2_navit.py:773-777, 824-829). Every layer keeps its own factors exactly as the reference does, although on
    a 'Taylor' step only the LAST layer's extrapolation reaches the output (each layer's result replaces its input)."""

    def __init__(self, num_layers: int, num_steps: int):
        self.factors: List[Dict[int, torch.Tensor]] = [dict() for _ in range(num_layers)]
        self.cache_counter = 0
        self.fresh_threshold = 3
        self.cal_threshold = None
        self.max_order = 6
        self.first_enhance = 5
        self.activated_steps = [0]
        self.step = 0
        self.num_steps = num_steps
        self.type = None

    def cal_type(self):
        first = self.step < self.first_enhance
        interval = self.fresh_threshold if first else self.cal_threshold
        if first or self.cache_counter == interval - 1:
            self.type = "full"
            self.cache_counter = 0
            self.activated_steps.append(self.step)
            self.cal_threshold = int(round(self.fresh_threshold / 1.0))   # force_scheduler, step_factor == 1
        else:
            self.cache_counter += 1
            self.type = "Taylor"

    def derivative_approximation(self, layer: int, feature: torch.Tensor):
        dist = self.activated_steps[-1] - self.activated_steps[-2]
        old = self.factors[layer]
        new = {0: feature}
        for i in range(self.max_order):
            if old.get(i, None) is not None and self.step > self.first_enhance - 2:
                new[i + 1] = (new[i] - old[i]) / dist
            else:
                break
        self.factors[layer] = new

    def taylor_formula(self, layer: int):
        import math
        x = self.step - self.activated_steps[-1]
        out = 0
        f = self.factors[layer]
        for i in range(len(f)):
            out += (1 / math.factorial(i)) * f[i] * (x ** i)
        return out


def _layer(x, sd, cfg, li, cos, sin, mode, vae_idx, text_idx, **attn_kw):
    """qwen2_navit.py:757-831 (Qwen2MoTDecoderLayer.forward_inference, TaylorSeer off)."""
    p = f"model.layers.{li}."
    eps = cfg.rms_norm_eps
    resid = x
    if cfg.layer_module != "Qwen2MoTDecoderLayer":
        # Qwen2DecoderLayer (:603-684) / Qwen2MoEDecoderLayer (:834-933): one set of norms + attention for every token;
        # the MoE layer routes only the MLP (text rows -> mlp, latent rows -> mlp_moe_gen) in mode "gen"
        h = rms_norm(x, sd[p + "input_layernorm.weight"], eps)
        a = _attention(h, sd, cfg, li, cos, sin, mode="und", vae_idx=vae_idx, text_idx=text_idx, **attn_kw)
        x = resid + a
        resid = x
        h = rms_norm(x, sd[p + "post_attention_layernorm.weight"], eps)
        if mode == "und" or cfg.layer_module == "Qwen2DecoderLayer":
            m = swiglu_mlp(h, sd, p + "mlp.")
        else:
            m = torch.zeros_like(h).to(BF16 if _AUTOCAST[0] == BF16 else h.dtype)
            m[text_idx] = swiglu_mlp(h[text_idx], sd, p + "mlp.")
            m[vae_idx] = swiglu_mlp(h[vae_idx], sd, p + "mlp_moe_gen.")
        return resid + m
    if mode == "und":
        h = rms_norm(x, sd[p + "input_layernorm.weight"], eps)
    else:
        h = torch.zeros_like(x)
        h[text_idx] = rms_norm(x[text_idx], sd[p + "input_layernorm.weight"], eps)
        h[vae_idx] = rms_norm(x[vae_idx], sd[p + "input_layernorm_moe_gen.weight"], eps)
    a = _attention(h, sd, cfg, li, cos, sin, mode=mode, vae_idx=vae_idx, text_idx=text_idx, **attn_kw)
    x = resid + a
    resid = x
    if mode == "und":
        h = rms_norm(x, sd[p + "post_attention_layernorm.weight"], eps)
        m = swiglu_mlp(h, sd, p + "mlp.")
    else:
        ht = rms_norm(x[text_idx], sd[p + "post_attention_layernorm.weight"], eps).to(_AUTOCAST[0])
        hv = rms_norm(x[vae_idx], sd[p + "post_attention_layernorm_moe_gen.weight"], eps).to(_AUTOCAST[0])
        m = torch.zeros_like(x).to(_AUTOCAST[0])
        m[text_idx] = swiglu_mlp(ht, sd, p + "mlp.")
        m[vae_idx] = swiglu_mlp(hv, sd, p + "mlp_moe_gen.")
    return resid + m


def lm_forward_inference(sd: Dict[str, torch.Tensor], cfg: LMConfig, packed_query_sequence, query_lens,
                         packed_query_position_ids, packed_query_indexes, past_key_values: Optional[KVCache] = None,
                         key_values_lens=None, packed_key_value_indexes=None, update_past_key_values=True,
                         is_causal=True, mode="und", packed_vae_token_indexes=None, packed_text_indexes=None,
                         taylor: Optional["TaylorSeerState"] = None):
    """qwen2_navit.py:1018-1092 (Qwen2Model.forward_inference). Returns (hidden [N,H], cache).
    `taylor`: TaylorSeer state of the calling branch (enable_taylorseer=True), else None."""
    x = packed_query_sequence
    if taylor is not None:
        taylor.cal_type()
    cos, sin = rope_tables(packed_query_position_ids, cfg.head_dim, cfg.rope_theta, x.dtype)
    for li in range(cfg.num_hidden_layers):
        if taylor is not None and taylor.type == "full" and taylor.step == 0:
            taylor.factors[li] = {}                                   # taylor_cache_init (:49-58)
        if taylor is None or taylor.type == "full":
            x = _layer(x, sd, cfg, li, cos, sin, mode, packed_vae_token_indexes, packed_text_indexes,
                       query_lens=query_lens, packed_query_indexes=packed_query_indexes, cache=past_key_values,
                       key_values_lens=key_values_lens, packed_key_value_indexes=packed_key_value_indexes,
                       update=update_past_key_values, is_causal=is_causal)
            if taylor is not None:
                taylor.derivative_approximation(li, x)
        else:
            x = taylor.taylor_formula(li)
    if taylor is not None:
        taylor.step += 1
    eps = cfg.rms_norm_eps
    if mode == "und" or not cfg.use_moe:
        x = rms_norm(x, sd["model.norm.weight"], eps)
    else:
        y = torch.zeros_like(x)
        y[packed_text_indexes] = rms_norm(x[packed_text_indexes], sd["model.norm.weight"], eps)
        y[packed_vae_token_indexes] = rms_norm(x[packed_vae_token_indexes], sd["model.norm_moe_gen.weight"], eps)
        x = y
    return x, past_key_values
