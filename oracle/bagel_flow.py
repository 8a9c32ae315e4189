"""TEST INFRASTRUCTURE — CPU restatement of BAGEL's packers, text prefill and rectified-flow sampler
(reference: modeling/bagel/bagel.py:232-297, 552-641, 644-907; modeling/bagel/modeling_utils.py:24-144;
data/data_utils.py:53-58). Pure functions over a state dict with the reference's parameter names
("language_model.*", "time_embedder.*", "vae2llm.*", "llm2vae.*", "latent_pos_embed.pos_embed").
Must NOT be run under torch.autocast: the autocast casts are written out explicitly (oracle.qwen2_mot.linear).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import qwen2_mot as om
from .qwen2_mot import BF16, KVCache, LMConfig, linear


@dataclass
class FlowConfig:
    lm: LMConfig
    latent_patch_size: int = 2
    latent_channel: int = 16
    vae_downsample: int = 8
    max_latent_size: int = 32

    @property
    def latent_downsample(self):
        return self.vae_downsample * self.latent_patch_size

    @property
    def patch_latent_dim(self):
        return self.latent_patch_size ** 2 * self.latent_channel


# ---------------------------------------------------------------------------------------------------
# tables / small heads
# ---------------------------------------------------------------------------------------------------
def sincos_1d(dim, pos):
    """modeling_utils.py:51-70"""
    omega = np.arange(dim // 2, dtype=np.float64)
    omega /= dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_2d_table(embed_dim, grid_size):
    """modeling_utils.py:24-48 (np.meshgrid(w, h): the w/column coordinate fills the first half)."""
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid_size, grid_size])
    emb = np.concatenate([sincos_1d(embed_dim // 2, grid[0]), sincos_1d(embed_dim // 2, grid[1])], axis=1)
    return torch.from_numpy(emb).float()


def timestep_embedding(t, dim=256, max_period=10000):
    """modeling_utils.py:87-105"""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def time_embedder(sd, t):
    """modeling_utils.py:107-110 under autocast: Linear -> SiLU -> Linear, bf16."""
    h = linear(timestep_embedding(t), sd["time_embedder.mlp.0.weight"], sd["time_embedder.mlp.0.bias"])
    return linear(F.silu(h), sd["time_embedder.mlp.2.weight"], sd["time_embedder.mlp.2.bias"])


def flattened_position_ids(img_h, img_w, patch, max_side):
    """data/data_utils.py:53-58"""
    nh, nw = img_h // patch, img_w // patch
    return (torch.arange(nh)[:, None] * max_side + torch.arange(nw)).flatten()


def lm_sub(sd):
    out = {k[len("language_model."):]: v for k, v in sd.items() if k.startswith("language_model.")}
    return out if type(sd) is dict else type(sd)(out)      # keeps om.LazyF32 lazy


# ---------------------------------------------------------------------------------------------------
# packers (bagel.py:232-264, 552-641)
# ---------------------------------------------------------------------------------------------------
def prepare_prompts(curr_kvlens, curr_rope, prompt_ids: List[List[int]], bos, eos):
    ids, pos, lens, t_idx, kv_idx = [], [], [], [], []
    cur = 0
    newlens, newrope = [], []
    for toks, kvlen, rope in zip(prompt_ids, curr_kvlens, curr_rope):
        kv_idx += list(range(cur, cur + kvlen))
        cur += kvlen
        toks = [bos] + list(toks) + [eos]
        lens.append(len(toks))
        ids += toks
        pos += list(range(rope, rope + len(toks)))
        t_idx += list(range(cur, cur + len(toks)))
        newlens.append(kvlen + len(toks))
        newrope.append(rope + len(toks))
        cur += len(toks)
    gi = {
        "text_token_lens": torch.tensor(lens, dtype=torch.int),
        "packed_text_ids": torch.tensor(ids, dtype=torch.long),
        "packed_text_position_ids": torch.tensor(pos, dtype=torch.long),
        "packed_text_indexes": torch.tensor(t_idx, dtype=torch.long),
        "packed_key_value_indexes": torch.tensor(kv_idx, dtype=torch.long),
        "key_values_lens": torch.tensor(curr_kvlens, dtype=torch.int),
    }
    return gi, newlens, newrope


def prepare_vae_latent(fc: FlowConfig, curr_kvlens, curr_rope, image_sizes, soi, eoi):
    t_ids, t_idx, vpos, v_idx, noises, pos, seqlens, q_idx, kv_idx = [], [], [], [], [], [], [], [], []
    qc = cur = 0
    for (H, W), kvlen, rope in zip(image_sizes, curr_kvlens, curr_rope):
        kv_idx += list(range(cur, cur + kvlen))
        cur += kvlen
        t_ids.append(soi); t_idx.append(qc); q_idx.append(cur); cur += 1; qc += 1
        vpos.append(flattened_position_ids(H, W, fc.latent_downsample, fc.max_latent_size))
        n = (H // fc.latent_downsample) * (W // fc.latent_downsample)
        noises.append(torch.randn(n, fc.patch_latent_dim))
        v_idx += list(range(qc, qc + n))
        q_idx += list(range(cur, cur + n))
        cur += n; qc += n
        t_ids.append(eoi); t_idx.append(qc); q_idx.append(cur); cur += 1; qc += 1
        pos += [rope] * (n + 2)
        seqlens.append(n + 2)
    return {
        "packed_text_ids": torch.tensor(t_ids, dtype=torch.long),
        "packed_text_indexes": torch.tensor(t_idx, dtype=torch.long),
        "packed_init_noises": torch.cat(noises, dim=0),
        "packed_vae_position_ids": torch.cat(vpos, dim=0),
        "packed_vae_token_indexes": torch.tensor(v_idx, dtype=torch.long),
        "packed_seqlens": torch.tensor(seqlens, dtype=torch.int),
        "packed_position_ids": torch.tensor(pos, dtype=torch.long),
        "key_values_lens": torch.tensor(curr_kvlens, dtype=torch.int),
        "packed_indexes": torch.tensor(q_idx, dtype=torch.long),
        "packed_key_value_indexes": torch.tensor(kv_idx, dtype=torch.long),
    }


def prepare_vae_latent_cfg(fc: FlowConfig, curr_kvlens, curr_rope, image_sizes):
    pos, q_idx, kv_idx = [], [], []
    cur = 0
    for (H, W), kvlen, rope in zip(image_sizes, curr_kvlens, curr_rope):
        kv_idx += list(range(cur, cur + kvlen))
        cur += kvlen
        n = (H // fc.latent_downsample) * (W // fc.latent_downsample)
        q_idx += list(range(cur, cur + n + 2))
        cur += n + 2
        pos += [rope] * (n + 2)
    return {
        "cfg_packed_position_ids": torch.tensor(pos, dtype=torch.long),
        "cfg_key_values_lens": torch.tensor(curr_kvlens, dtype=torch.int),
        "cfg_packed_query_indexes": torch.tensor(q_idx, dtype=torch.long),
        "cfg_packed_key_value_indexes": torch.tensor(kv_idx, dtype=torch.long),
    }


# ---------------------------------------------------------------------------------------------------
# prefill + sampler
# ---------------------------------------------------------------------------------------------------
def forward_cache_update_text(sd, fc: FlowConfig, cache: KVCache, packed_text_ids, packed_text_position_ids,
                              text_token_lens, packed_text_indexes, packed_key_value_indexes, key_values_lens):
    """bagel.py:267-297"""
    emb = F.embedding(packed_text_ids, sd["language_model.model.embed_tokens.weight"])
    _, cache = om.lm_forward_inference(lm_sub(sd), fc.lm, emb, text_token_lens, packed_text_position_ids,
                                       packed_text_indexes, cache, key_values_lens, packed_key_value_indexes,
                                       True, True, "und")
    return cache


def forward_flow(sd, fc: FlowConfig, x_t, timestep, packed_vae_token_indexes, packed_vae_position_ids,
                 packed_text_ids, packed_text_indexes, packed_indexes, packed_position_ids, packed_seqlens,
                 key_values_lens, past_key_values, packed_key_value_indexes, cfg_renorm_min=0.0,
                 cfg_renorm_type="global", cfg_text_scale=1.0, cfg_text=None, cfg_img_scale=1.0, cfg_img=None,
                 taylor=None):
    """bagel.py:757-907. cfg_text / cfg_img: dicts with packed_position_ids, packed_query_indexes,
    key_values_lens, past_key_values, packed_key_value_indexes of the respective branch.
    taylor: None or (main, text, img) TaylorSeerState triple (bagel.py:680-684, 816-818, 836-838, 855-857)."""
    lsd = lm_sub(sd)
    H = fc.lm.hidden_size
    emb = F.embedding(packed_text_ids, sd["language_model.model.embed_tokens.weight"])
    seq = emb.new_zeros((int(sum(packed_seqlens)), H))
    seq[packed_text_indexes] = emb
    assert timestep.unique().shape[0] == 1
    pos_embed = sd["latent_pos_embed.pos_embed"][packed_vae_position_ids]
    t_emb = time_embedder(sd, timestep)
    lat = linear(x_t, sd["vae2llm.weight"], sd["vae2llm.bias"]) + t_emb + pos_embed
    if lat.dtype != seq.dtype:
        lat = lat.to(seq.dtype)
    seq[packed_vae_token_indexes] = lat

    def branch(pos_ids, q_idx, kv_lens, cache, kv_idx, ts_state=None):
        h, _ = om.lm_forward_inference(lsd, fc.lm, seq, packed_seqlens, pos_ids, q_idx, cache, kv_lens, kv_idx,
                                       False, False, "gen", packed_vae_token_indexes, packed_text_indexes,
                                       taylor=ts_state)
        return linear(h, sd["llm2vae.weight"], sd["llm2vae.bias"])[packed_vae_token_indexes]

    ts_main, ts_text, ts_img = taylor if taylor is not None else (None, None, None)
    v = branch(packed_position_ids, packed_indexes, key_values_lens, past_key_values, packed_key_value_indexes, ts_main)
    if cfg_text_scale > 1.0:
        vT = branch(cfg_text["packed_position_ids"], cfg_text["packed_query_indexes"], cfg_text["key_values_lens"],
                    cfg_text["past_key_values"], cfg_text["packed_key_value_indexes"], ts_text)
    if cfg_img_scale > 1.0:
        vI = branch(cfg_img["packed_position_ids"], cfg_img["packed_query_indexes"], cfg_img["key_values_lens"],
                    cfg_img["past_key_values"], cfg_img["packed_key_value_indexes"], ts_img)
    if cfg_text_scale > 1.0:
        u = vT + cfg_text_scale * (v - vT)
        if cfg_renorm_type == "text_channel":
            sc = (torch.norm(v, dim=-1, keepdim=True) / (torch.norm(u, dim=-1, keepdim=True) + 1e-8)).clamp(
                min=cfg_renorm_min, max=1.0)
            ut = u * sc
            v = vI + cfg_img_scale * (ut - vI) if cfg_img_scale > 1.0 else ut
        else:
            w = vI + cfg_img_scale * (u - vI) if cfg_img_scale > 1.0 else u
            if cfg_renorm_type == "global":
                nv, nw = torch.norm(v), torch.norm(w)
            elif cfg_renorm_type == "channel":
                nv, nw = torch.norm(v, dim=-1, keepdim=True), torch.norm(w, dim=-1, keepdim=True)
            else:
                raise NotImplementedError(cfg_renorm_type)
            v = w * (nv / (nw + 1e-8)).clamp(min=cfg_renorm_min, max=1.0)
    return v


def generate_image(sd, fc: FlowConfig, gen_input: Dict, past_key_values, num_timesteps=24, timestep_shift=1.0,
                   cfg_renorm_min=0.0, cfg_renorm_type="global", cfg_interval=(0, 1), cfg_text_scale=1.0,
                   cfg_text=None, cfg_img_scale=1.0, cfg_img=None, trace: Optional[list] = None,
                   enable_taylorseer: bool = False, x_trace: Optional[list] = None, max_steps: Optional[int] = None):
    """bagel.py:644-754. gen_input = prepare_vae_latent(...) dict. Returns the tuple of per-sample latents.
    enable_taylorseer: one TaylorSeer cache per branch, created per call (bagel.py:680-684)."""
    x_t = gen_input["packed_init_noises"]
    taylor = None
    if enable_taylorseer:
        taylor = tuple(om.TaylorSeerState(fc.lm.num_hidden_layers, num_timesteps) for _ in range(3))
    ts = torch.linspace(1, 0, num_timesteps)
    ts = timestep_shift * ts / (1 + (timestep_shift - 1) * ts)
    dts = ts[:-1] - ts[1:]
    ts = ts[:-1]
    for i, t in enumerate(ts):
        if max_steps is not None and i >= max_steps:
            break
        timestep = torch.tensor([t] * x_t.shape[0]).to(x_t.device)
        on = bool(t > cfg_interval[0] and t <= cfg_interval[1])
        v = forward_flow(sd, fc, x_t, timestep, gen_input["packed_vae_token_indexes"],
                         gen_input["packed_vae_position_ids"], gen_input["packed_text_ids"],
                         gen_input["packed_text_indexes"], gen_input["packed_indexes"],
                         gen_input["packed_position_ids"], gen_input["packed_seqlens"], gen_input["key_values_lens"],
                         past_key_values, gen_input["packed_key_value_indexes"], cfg_renorm_min, cfg_renorm_type,
                         cfg_text_scale if on else 1.0, cfg_text, cfg_img_scale if on else 1.0, cfg_img, taylor=taylor)
        if trace is not None:
            trace.append(v.clone())
        x_t = x_t - v * dts[i].to(x_t.device)
        if x_trace is not None:
            x_trace.append(x_t.clone())
    return x_t.split((gen_input["packed_seqlens"] - 2).tolist())


def prepare_start_tokens(curr_kvlens, curr_rope, bos):
    """bagel.py:909-927"""
    kv_idx, cur = [], 0
    for kvlen in curr_kvlens:
        kv_idx += list(range(cur, cur + kvlen))
        cur += kvlen
    return {
        "packed_start_tokens": torch.tensor([bos] * len(curr_kvlens), dtype=torch.long),
        "packed_query_position_ids": torch.tensor(list(curr_rope), dtype=torch.long),
        "key_values_lens": torch.tensor(curr_kvlens, dtype=torch.int),
        "packed_key_value_indexes": torch.tensor(kv_idx, dtype=torch.long),
    }


def generate_text(sd, fc: FlowConfig, past_key_values, packed_key_value_indexes, key_values_lens,
                  packed_start_tokens, packed_query_position_ids, max_length, end_token_id=None,
                  logits_trace: Optional[list] = None):
    """Greedy branch of bagel.py:930-1000 (do_sample=False). Returns [steps, B] ids."""
    lsd = lm_sub(sd)
    seq, cur = [], packed_start_tokens
    kv = key_values_lens.clone()
    pos = packed_query_position_ids.clone()
    kv_idx = packed_key_value_indexes.clone()
    for _ in range(max_length):
        seq.append(cur)
        emb = F.embedding(cur, sd["language_model.model.embed_tokens.weight"])
        q_idx = torch.cumsum(kv, dim=0) + torch.arange(0, len(kv), dtype=kv.dtype)
        parts = list(kv_idx.split(kv.tolist(), dim=0))
        kv_idx = torch.cat([p + i for i, p in enumerate(parts)], dim=0)
        h, past_key_values = om.lm_forward_inference(lsd, fc.lm, emb, torch.ones_like(cur), pos, q_idx, past_key_values,
                                                     kv, kv_idx, True, True, "und")
        logits = linear(h, sd["language_model.lm_head.weight"])
        if logits_trace is not None:
            logits_trace.append(logits.clone())
        cur = torch.argmax(logits, dim=-1)
        parts = list(kv_idx.split(kv.tolist(), dim=0))
        kv_idx = torch.cat([torch.cat([p, p[-1:] + 1]) for p in parts], dim=0)
        kv = kv + 1
        pos = pos + 1
        if end_token_id is not None and cur[0] == end_token_id:
            break
    return torch.stack(seq, dim=0)


def prepare_vae_images(fc: FlowConfig, curr_kvlens, curr_rope, image_tensors, soi, eoi, timestep=0):
    """bagel.py:417-488 with `transforms` already applied."""
    shapes, vpos, v_idx, t_ids, t_idx, seqlens, pos, q_idx, kv_idx = [], [], [], [], [], [], [], [], []
    qc = cur = 0
    newlens, newrope = [], []
    for img, kvlen, rope in zip(image_tensors, curr_kvlens, curr_rope):
        kv_idx += list(range(cur, cur + kvlen)); cur += kvlen
        t_ids.append(soi); t_idx.append(qc); q_idx.append(cur); cur += 1; qc += 1
        vpos.append(flattened_position_ids(img.size(1), img.size(2), fc.latent_downsample, fc.max_latent_size))
        h, w = img.shape[1] // fc.latent_downsample, img.shape[2] // fc.latent_downsample
        shapes.append((h, w)); n = h * w
        v_idx += list(range(qc, qc + n)); q_idx += list(range(cur, cur + n)); cur += n; qc += n
        t_ids.append(eoi); t_idx.append(qc); q_idx.append(cur); cur += 1; qc += 1
        pos += [rope] * (n + 2); seqlens.append(n + 2)
        newlens.append(kvlen + n + 2); newrope.append(rope + 1)
    sizes = [t.shape for t in image_tensors]
    mx = [max(v) for v in zip(*sizes)]
    padded = torch.zeros((len(image_tensors), *mx))
    for i, t in enumerate(image_tensors):
        padded[i, :, : t.shape[1], : t.shape[2]] = t
    gi = {
        "padded_images": padded, "patchified_vae_latent_shapes": shapes,
        "packed_vae_position_ids": torch.cat(vpos, dim=0), "packed_timesteps": torch.tensor([timestep]),
        "packed_vae_token_indexes": torch.tensor(v_idx, dtype=torch.long),
        "packed_text_ids": torch.tensor(t_ids, dtype=torch.long),
        "packed_text_indexes": torch.tensor(t_idx, dtype=torch.long),
        "packed_position_ids": torch.tensor(pos, dtype=torch.long),
        "packed_seqlens": torch.tensor(seqlens, dtype=torch.int),
        "packed_indexes": torch.tensor(q_idx, dtype=torch.long),
        "packed_key_value_indexes": torch.tensor(kv_idx, dtype=torch.long),
        "key_values_lens": torch.tensor(curr_kvlens, dtype=torch.int),
    }
    return gi, newlens, newrope


def forward_cache_update_vae(sd, fc: FlowConfig, vae_encode, cache, padded_images, patchified_vae_latent_shapes,
                             packed_vae_position_ids, packed_timesteps, packed_vae_token_indexes, packed_text_ids,
                             packed_text_indexes, packed_position_ids, packed_seqlens, packed_indexes, key_values_lens,
                             packed_key_value_indexes):
    """bagel.py:491-550; `vae_encode(images) -> latents [B, z, H/8, W/8]` stands for vae_model.encode."""
    emb = F.embedding(packed_text_ids, sd["language_model.model.embed_tokens.weight"])
    seq = emb.new_zeros((int(sum(packed_seqlens)), fc.lm.hidden_size))
    seq[packed_text_indexes] = emb
    lat = vae_encode(padded_images)
    p = fc.latent_patch_size
    rows = []
    for z, (h, w) in zip(lat, patchified_vae_latent_shapes):
        z = z[:, : h * p, : w * p].reshape(fc.latent_channel, h, p, w, p)
        rows.append(torch.einsum("chpwq->hwpqc", z).reshape(-1, p * p * fc.latent_channel))
    packed = torch.cat(rows, dim=0)
    packed = linear(packed, sd["vae2llm.weight"], sd["vae2llm.bias"]) + time_embedder(sd, packed_timesteps) \
        + sd["latent_pos_embed.pos_embed"][packed_vae_position_ids]
    if packed.dtype != seq.dtype:
        packed = packed.to(seq.dtype)
    seq[packed_vae_token_indexes] = packed
    _, cache = om.lm_forward_inference(lm_sub(sd), fc.lm, seq, packed_seqlens, packed_position_ids, packed_indexes,
                                       cache, key_values_lens, packed_key_value_indexes, True, False, "gen",
                                       packed_vae_token_indexes, packed_text_indexes)
    return cache
