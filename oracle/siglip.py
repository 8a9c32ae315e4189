"""TEST INFRASTRUCTURE — CPU restatement of the SigLIP NaViT encoder and the ViT context prefill
(reference: modeling/bagel/siglip_navit.py:145-402 with rope=False; modeling/bagel/bagel.py:299-415;
data/data_utils.py:43-50). State-dict keys as in the reference ("vit_model.vision_model.*", "connector.*",
"vit_pos_embed.pos_embed"). Written for mode A (bf16 parameters under CPU autocast): nn.Linear runs in bf16,
nn.LayerNorm is not on the CPU autocast list and therefore runs in the dtype of its input."""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import qwen2_mot as om
from .bagel_flow import flattened_position_ids, lm_sub
from .qwen2_mot import linear


@dataclass
class VitConfig:
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    patch_size: int = 14
    num_channels: int = 3
    layer_norm_eps: float = 1e-6
    rope: bool = False          # 2-D RoPE on q/k (siglip_navit.py:102-142, 224-230); off in every shipped loader
    image_size: int = 980       # RoPE table side = image_size // patch_size (:343-345)


def patchify(image, p):
    """data/data_utils.py:43-50"""
    c, h, w = image.shape
    image = image.reshape(c, h // p, p, w // p, p)
    return torch.einsum("chpwq->hwpqc", image).reshape(-1, p * p * c)


def layer_norm(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w.to(x.dtype), b.to(x.dtype), eps)


def rope2d_tables(dim, max_h, max_w, base=10000):
    """RotaryEmbedding2D (siglip_navit.py:102-127): fp32 buffers cos_h, sin_h, cos_w, sin_w of shape [max_h*max_w, dim]."""
    inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))
    grid_h = torch.arange(0, max_h).to(inv_freq.dtype)[:, None].repeat(1, max_w)
    grid_w = torch.arange(0, max_w).to(inv_freq.dtype)[None, :].repeat(max_h, 1)

    def side(grid):
        freqs = grid[..., None] * inv_freq[None, None, :]
        emb = torch.cat((freqs, freqs), dim=-1).flatten(0, 1)
        return emb.cos(), emb.sin()

    return (*side(grid_h), *side(grid_w))


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def vit_forward(sd, vc: VitConfig, pixels, pos_ids, seqlens, pfx="vit_model.vision_model."):
    """siglip_navit.py:184-195, 216-243, 255-258, 283-298, 354-371."""
    if vc.rope:
        side = vc.image_size // vc.patch_size
        cos_h, sin_h, cos_w, sin_w = (t[pos_ids].unsqueeze(1) for t in
                                      rope2d_tables(vc.hidden_size // vc.num_attention_heads // 2, side, side))
    x = linear(pixels, sd[pfx + "embeddings.patch_embedding.weight"], sd[pfx + "embeddings.patch_embedding.bias"])
    if not vc.rope:     # with rope=True the tower has no learned position table at all (:164-165, 191-194)
        x = x + sd[pfx + "embeddings.position_embedding.weight"][pos_ids]
    nh = vc.num_attention_heads
    d = vc.hidden_size // nh
    lens = [int(v) for v in seqlens]
    for li in range(vc.num_hidden_layers):
        p = pfx + f"encoder.layers.{li}."
        h = layer_norm(x, sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], vc.layer_norm_eps)
        q = linear(h, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"]).view(-1, nh, d)
        k = linear(h, sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"]).view(-1, nh, d)
        v = linear(h, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"]).view(-1, nh, d)
        if vc.rope:     # :224-230 — halves of head_dim rotated by the row / column tables (fp32 buffers => fp32 math)
            qh, qw, kh, kw = q[..., : d // 2], q[..., d // 2:], k[..., : d // 2], k[..., d // 2:]
            qh, kh = qh * cos_h + _rot_half(qh) * sin_h, kh * cos_h + _rot_half(kh) * sin_h
            qw, kw = qw * cos_w + _rot_half(qw) * sin_w, kw * cos_w + _rot_half(kw) * sin_w
            q, k = torch.cat([qh, qw], dim=-1), torch.cat([kh, kw], dim=-1)
        dt = om._AUTOCAST[0]
        a = om.varlen_attention(q.to(dt), k.to(dt), v.to(dt), lens, lens, False).reshape(-1, nh * d)
        x = x + linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
        h = layer_norm(x, sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], vc.layer_norm_eps)
        h = linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
        h = F.gelu(h, approximate="tanh")
        x = x + linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return layer_norm(x, sd[pfx + "post_layernorm.weight"], sd[pfx + "post_layernorm.bias"], vc.layer_norm_eps)


def connector(sd, x):
    """modeling_utils.py:120-124"""
    h = linear(x, sd["connector.fc1.weight"], sd["connector.fc1.bias"])
    return linear(F.gelu(h, approximate="tanh"), sd["connector.fc2.weight"], sd["connector.fc2.bias"])


def prepare_vit_images(vc: VitConfig, max_side, curr_kvlens, curr_rope, image_tensors, soi, eoi):
    """bagel.py:299-359 with `transforms` already applied (image_tensors are [C,H,W] floats)."""
    t_ids, t_idx, v_idx, toks, vpos, lens, pos, seqlens, q_idx, kv_idx = [], [], [], [], [], [], [], [], [], []
    qc = cur = 0
    newlens, newrope = [], []
    for img, kvlen, rope in zip(image_tensors, curr_kvlens, curr_rope):
        kv_idx += list(range(cur, cur + kvlen)); cur += kvlen
        t_ids.append(soi); t_idx.append(qc); q_idx.append(cur); cur += 1; qc += 1
        vpos.append(flattened_position_ids(img.size(1), img.size(2), vc.patch_size, max_side))
        tk = patchify(img, vc.patch_size)
        toks.append(tk); n = tk.shape[0]; lens.append(n)
        v_idx += list(range(qc, qc + n)); q_idx += list(range(cur, cur + n)); cur += n; qc += n
        t_ids.append(eoi); t_idx.append(qc); q_idx.append(cur); cur += 1; qc += 1
        pos += [rope] * (n + 2); seqlens.append(n + 2)
        newlens.append(kvlen + n + 2); newrope.append(rope + 1)
    gi = {
        "packed_text_ids": torch.tensor(t_ids, dtype=torch.long),
        "packed_text_indexes": torch.tensor(t_idx, dtype=torch.long),
        "vit_token_seqlens": torch.tensor(lens, dtype=torch.int),
        "packed_vit_tokens": torch.cat(toks, dim=0),
        "packed_vit_position_ids": torch.cat(vpos, dim=0),
        "packed_vit_token_indexes": torch.tensor(v_idx, dtype=torch.long),
        "packed_position_ids": torch.tensor(pos, dtype=torch.long),
        "packed_seqlens": torch.tensor(seqlens, dtype=torch.int),
        "packed_indexes": torch.tensor(q_idx, dtype=torch.long),
        "packed_key_value_indexes": torch.tensor(kv_idx, dtype=torch.long),
        "key_values_lens": torch.tensor(curr_kvlens, dtype=torch.int),
    }
    return gi, newlens, newrope


def forward_cache_update_vit(sd, lm_cfg, vc: VitConfig, cache, packed_text_ids, packed_text_indexes, packed_vit_tokens,
                             packed_vit_token_indexes, packed_vit_position_ids, vit_token_seqlens,
                             packed_position_ids, packed_seqlens, packed_indexes, packed_key_value_indexes,
                             key_values_lens):
    """bagel.py:362-415"""
    emb = F.embedding(packed_text_ids, sd["language_model.model.embed_tokens.weight"])
    seq = emb.new_zeros((int(sum(packed_seqlens)), lm_cfg.hidden_size))
    seq[packed_text_indexes] = emb
    feats = vit_forward(sd, vc, packed_vit_tokens, packed_vit_position_ids, vit_token_seqlens)
    feats = connector(sd, feats) + sd["vit_pos_embed.pos_embed"][packed_vit_position_ids]
    if feats.dtype != seq.dtype:
        feats = feats.to(seq.dtype)
    seq[packed_vit_token_indexes] = feats
    _, cache = om.lm_forward_inference(lm_sub(sd), lm_cfg, seq, packed_seqlens, packed_position_ids, packed_indexes,
                                       cache, key_values_lens, packed_key_value_indexes, True, False, "und")
    return cache
