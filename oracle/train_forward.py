"""TEST INFRASTRUCTURE — CPU restatement of the TRAINING-mode forward (no backward) of BAGEL:
`Bagel.forward` (modeling/bagel/bagel.py:101-229), `Qwen2Model.forward_train` (qwen2_navit.py:970-1016),
`Qwen2MoTDecoderLayer.forward_train` (:713-755), `PackedAttentionMoT.forward_train` (:406-497, the
`nested_attention_masks` branch: per-sample SDPA with a dense additive mask) and the mask algebra
`prepare_attention_mask_per_sample` (data/data_utils.py:72-103: splits of a sample are 'causal', 'full' or 'noise').
Pinned bit-for-bit against the reference by tests/golden/make_golden.py (golden_train_forward).

Dtype flow under autocast with bf16 parameters: unlike the inference gen branch there is NO fp32 upcast around q/k-norm and
RoPE here — every token (und and gen) takes the all-bf16 flow; only the expert weights differ.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import bagel_flow as obf
from . import qwen2_mot as om
from .qwen2_mot import LMConfig, apply_rope, linear, rms_norm, rope_tables, swiglu_mlp


def prepare_attention_mask_per_sample(split_lens, attn_modes):
    """data/data_utils.py:72-103 — additive fp32 mask [L, L] (0 = attend, -inf = ignore) of one sample."""
    n = sum(split_lens)
    m = torch.zeros((n, n), dtype=torch.bool)
    c = 0
    for s, mode in zip(split_lens, attn_modes):
        assert mode in ("causal", "full", "noise")
        m[c:c + s, c:c + s] = torch.ones((s, s)).tril() if mode == "causal" else torch.ones((s, s))
        m[c:c + s, :c] = 1
        c += s
    c = 0
    for s, mode in zip(split_lens, attn_modes):
        if mode == "noise":          # nobody but the split itself sees a noised image
            m[:, c:c + s] = torch.zeros((n, s))
            m[c:c + s, c:c + s] = torch.ones((s, s))
        c += s
    return torch.zeros_like(m, dtype=torch.float).masked_fill_(~m, float("-inf"))


def _masked_attention(q, k, v, sample_lens, masks, Hq, Hk):
    """qwen2_navit.py:451-473: GQA by repeating K/V heads, per-sample scaled_dot_product_attention with the dense mask
    (the same torch op the reference calls; bf16 operands under autocast, fp32 in high_precision mode)."""
    dt = om._AUTOCAST[0]
    g = Hq // Hk
    k = k[:, :, None, :].repeat(1, 1, g, 1).reshape(-1, Hq, k.shape[-1])
    v = v[:, :, None, :].repeat(1, 1, g, 1).reshape(-1, Hq, v.shape[-1])
    outs = []
    for qs, ks, vs, mask in zip(q.transpose(0, 1).split(sample_lens, dim=1), k.transpose(0, 1).split(sample_lens, dim=1),
                                v.transpose(0, 1).split(sample_lens, dim=1), masks):
        # the reference wraps this call in sdpa_kernel([EFFICIENT_ATTENTION]) (:462), a CUDA backend selector with no viable
        # CPU kernel; the golden harness replaces that context manager by a no-op (oracle/ref_shims.py, shim 4)
        o = F.scaled_dot_product_attention(qs.to(dt).unsqueeze(0), ks.to(dt).unsqueeze(0), vs.to(dt).unsqueeze(0),
                                           mask.to(dt).unsqueeze(0))
        outs.append(o.squeeze(0))
    return torch.cat(outs, dim=1).transpose(0, 1)          # [N, Hq, d]


def _attention_train(x, sd, cfg: LMConfig, li, cos, sin, sample_lens, masks, und, gen):
    p = f"model.layers.{li}.self_attn."
    Hq, Hk, d = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    eps = cfg.rms_norm_eps
    n = x.shape[0]
    q, k, v = x.new_zeros((n, Hq * d)), x.new_zeros((n, Hk * d)), x.new_zeros((n, Hk * d))
    xu, xg = x[und], x[gen]
    for dst, name in ((q, "q"), (k, "k"), (v, "v")):
        dst[und] = linear(xu, sd[p + f"{name}_proj.weight"], sd[p + f"{name}_proj.bias"])
        dst[gen] = linear(xg, sd[p + f"{name}_proj_moe_gen.weight"], sd[p + f"{name}_proj_moe_gen.bias"])
    q, k, v = q.view(-1, Hq, d), k.view(-1, Hk, d), v.view(-1, Hk, d)
    q_, k_ = q.new_zeros(q.shape), k.new_zeros(k.shape)
    q_[und] = rms_norm(q[und], sd[p + "q_norm.weight"], eps)
    q_[gen] = rms_norm(q[gen], sd[p + "q_norm_moe_gen.weight"], eps)
    k_[und] = rms_norm(k[und], sd[p + "k_norm.weight"], eps)
    k_[gen] = rms_norm(k[gen], sd[p + "k_norm_moe_gen.weight"], eps)
    q_, k_ = apply_rope(q_, k_, cos, sin)
    o = _masked_attention(q_, k_, v, sample_lens, masks, Hq, Hk).reshape(-1, Hq * d)
    out = o.new_zeros(o.shape)
    out[und] = linear(o[und], sd[p + "o_proj.weight"])
    out[gen] = linear(o[gen], sd[p + "o_proj_moe_gen.weight"])
    return out


def lm_forward_train(sd, cfg: LMConfig, x, sample_lens, masks, position_ids, und, gen):
    """Qwen2Model.forward_train (MoT layers). Returns the final-normed hidden states [N, H]."""
    eps = cfg.rms_norm_eps
    cos, sin = rope_tables(position_ids, cfg.head_dim, cfg.rope_theta, x.dtype)
    for li in range(cfg.num_hidden_layers):
        p = f"model.layers.{li}."
        resid = x
        h = x.new_zeros(x.shape)
        h[und] = rms_norm(x[und], sd[p + "input_layernorm.weight"], eps)
        h[gen] = rms_norm(x[gen], sd[p + "input_layernorm_moe_gen.weight"], eps)
        x = resid + _attention_train(h, sd, cfg, li, cos, sin, sample_lens, masks, und, gen)
        resid = x
        m = x.new_zeros(x.shape)
        m[und] = swiglu_mlp(rms_norm(x[und], sd[p + "post_attention_layernorm.weight"], eps), sd, p + "mlp.")
        m[gen] = swiglu_mlp(rms_norm(x[gen], sd[p + "post_attention_layernorm_moe_gen.weight"], eps), sd, p + "mlp_moe_gen.")
        x = resid + m
    y = torch.zeros_like(x)
    y[und] = rms_norm(x[und], sd["model.norm.weight"], eps)
    y[gen] = rms_norm(x[gen], sd["model.norm_moe_gen.weight"], eps)
    return y


def bagel_forward_train(sd, fc: obf.FlowConfig, sequence_length, packed_text_ids, packed_text_indexes, sample_lens,
                        packed_position_ids, nested_attention_masks, noise, timestep_shift=1.0, ce_loss_indexes=None,
                        packed_label_ids=None, vit=None, padded_latent=None, patchified_vae_latent_shapes=None,
                        packed_latent_position_ids=None, packed_vae_token_indexes=None, packed_timesteps=None,
                        mse_loss_indexes=None) -> Dict[str, Optional[torch.Tensor]]:
    """Bagel.forward with `nested_attention_masks`. `noise` replaces the torch.randn_like draw (bagel.py:184);
    `vit` = None or (VitConfig, packed_vit_tokens, packed_vit_token_indexes, packed_vit_position_ids, vit_token_seqlens)."""
    from . import siglip as osl
    lsd = obf.lm_sub(sd)
    H = fc.lm.hidden_size
    emb = F.embedding(packed_text_ids, sd["language_model.model.embed_tokens.weight"])
    seq = emb.new_zeros((sequence_length, H))
    seq[packed_text_indexes] = emb
    und = packed_text_indexes
    if vit is not None:
        vc, vit_tokens, vit_idx, vit_pos, vit_lens = vit
        feats = osl.connector(sd, osl.vit_forward(sd, vc, vit_tokens, vit_pos, vit_lens))
        feats = feats + sd["vit_pos_embed.pos_embed"][vit_pos]
        seq[vit_idx] = feats.to(seq.dtype)
        und = torch.cat([packed_text_indexes, vit_idx], dim=0)
    p = fc.latent_patch_size
    rows = []
    for lat, (h, w) in zip(padded_latent, patchified_vae_latent_shapes):
        lat = lat[:, : h * p, : w * p].reshape(fc.latent_channel, h, p, w, p)
        rows.append(torch.einsum("chpwq->hwpqc", lat).reshape(-1, p * p * fc.latent_channel))
    clean = torch.cat(rows, dim=0)
    t = torch.sigmoid(packed_timesteps)
    t = timestep_shift * t / (1 + (timestep_shift - 1) * t)
    x_t = (1 - t[:, None]) * clean + t[:, None] * noise
    lat = linear(x_t, sd["vae2llm.weight"], sd["vae2llm.bias"]) + obf.time_embedder(sd, t) \
        + sd["latent_pos_embed.pos_embed"][packed_latent_position_ids]
    seq[packed_vae_token_indexes] = lat.to(seq.dtype)
    hidden = lm_forward_train(lsd, fc.lm, seq, list(sample_lens), nested_attention_masks, packed_position_ids, und,
                              packed_vae_token_indexes)
    preds = linear(hidden[mse_loss_indexes], sd["llm2vae.weight"], sd["llm2vae.bias"])
    target = noise - clean
    mse = (preds - target[t > 0]) ** 2
    ce = None
    if ce_loss_indexes is not None:
        logits = linear(hidden[ce_loss_indexes], sd["language_model.lm_head.weight"])
        ce = F.cross_entropy(logits.float(), packed_label_ids, reduction="none")
    return dict(mse=mse, ce=ce, last_hidden_state=hidden)
