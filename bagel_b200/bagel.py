"""BAGEL unified model — host side of the B200-native build.

Same public surface as the reference's modeling/bagel/bagel.py (Bagel :57): the `prepare_*` packers
(:232-264, :552-641, :909-927) return dicts with the same keys / dtypes (key names are API — callers splat
them as kwargs), `forward_cache_update_text` (:267-297) prefill, the rectified-flow sampler `generate_image`
(:644-754) with `_forward_flow` (:757-907), and `generate_text` (:930-1000).

What is different underneath (B200-first, not a port):
  * packers are vectorised index arithmetic (no per-token Python loops);
  * `generate_image` plans the whole run once (index maps, RoPE tables, all timestep embeddings, merged KV
    buffers with the read-only context already in place), batches the CFG branches into ONE packed LM call per
    step (they share weights; the reference runs them one after another), keeps x_t resident in HBM and issues
    a sync-free launch sequence per step, optionally replayed as a CUDA graph;
  * every tensor op on the per-step path is a kernel from bagel_b200.ops.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .config import BagelConfig
from .modeling_utils import MLPconnector, PositionEmbedding, TimestepEmbedder
from .qwen2_navit import NaiveCache, Qwen2ForCausalLM

BF16 = torch.bfloat16


# --------------------------------------------------------------------------------------------------
# index helpers (host, int64)
# --------------------------------------------------------------------------------------------------
def _ranges(starts: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
    """concat_i arange(starts[i], starts[i] + lens[i])"""
    lens = lens.to(torch.int64)
    total = int(lens.sum())
    if total == 0:
        return torch.zeros(0, dtype=torch.int64)
    excl = torch.cumsum(lens, 0) - lens
    return torch.repeat_interleave(starts.to(torch.int64) - excl, lens) + torch.arange(total, dtype=torch.int64)


def get_flattened_position_ids_extrapolate(img_h, img_w, patch_size, max_num_patches_per_side):
    """row * max_side + col for every patch (reference data/data_utils.py:53-58)."""
    nh, nw = img_h // patch_size, img_w // patch_size
    return (torch.arange(nh)[:, None] * max_num_patches_per_side + torch.arange(nw)[None, :]).reshape(-1)


def get_flattened_position_ids_interpolate(img_h, img_w, patch_size, max_num_patches_per_side):
    """Bucketised fractional coordinates (reference data/data_utils.py:61-69)."""
    nh, nw = img_h // patch_size, img_w // patch_size
    edges = torch.arange(1 / max_num_patches_per_side, 1.0, 1 / max_num_patches_per_side)
    bh = torch.bucketize(torch.arange(0, 1 - 1e-6, 1 / nh), edges, right=True)
    bw = torch.bucketize(torch.arange(0, 1 - 1e-6, 1 / nw), edges, right=True)
    return (bh[:, None] * max_num_patches_per_side + bw[None, :]).reshape(-1)


def patchify(image: torch.Tensor, patch_size: int) -> torch.Tensor:
    """[C, H, W] -> [(H/p)*(W/p), p*p*C], each patch flattened (row-in-patch, col-in-patch, channel) — the
    order convert_conv2d_to_linear's W.permute(0,2,3,1) expects (reference data/data_utils.py:43-50)."""
    c, h, w = image.shape
    p = patch_size
    assert h % p == 0 and w % p == 0
    return image.reshape(c, h // p, p, w // p, p).permute(1, 3, 2, 4, 0).reshape(-1, p * p * c)


class _Affine:
    """weight/bias holder for vae2llm / llm2vae."""

    def __init__(self):
        self.weight = self.bias = None

    def load(self, sd, prefix, device):
        self.weight = sd[prefix + "weight"].to(device, BF16).contiguous()
        self.bias = sd[prefix + "bias"].to(device, BF16).contiguous()

    def __call__(self, x):
        return ops.gemm(x.to(BF16).contiguous(), self.weight, bias=self.bias)


class Bagel:
    def __init__(self, language_model: Qwen2ForCausalLM, vit_model, config: BagelConfig):
        self.language_model = language_model
        self.config = config
        self.device = language_model.device
        # "A": bf16 weights + autocast (app.py:111); "B": fp32 master weights + autocast (eval drivers) — see qwen2_navit.py
        self.dtype_mode = language_model.model.dtype_mode
        sdt = language_model.model.stream_dtype
        llm = config.llm_config
        self.hidden_size = llm.hidden_size
        self.use_moe = "Mo" in llm.layer_module
        self.num_heads = llm.num_attention_heads
        if config.visual_gen:
            self.latent_patch_size = config.latent_patch_size
            self.timestep_shift = config.timestep_shift
            self.latent_downsample = config.vae_config.downsample * config.latent_patch_size
            self.max_latent_size = config.max_latent_size
            self.latent_channel = config.vae_config.z_channels
            self.patch_latent_dim = self.latent_patch_size ** 2 * self.latent_channel
            self.time_embedder = TimestepEmbedder(self.hidden_size)
            self.vae2llm = _Affine()
            self.llm2vae = _Affine()
            self.latent_pos_embed = PositionEmbedding(self.max_latent_size, self.hidden_size, self.device, sdt)
        if config.visual_und:
            self.vit_model = vit_model
            self.vit_patch_size = config.vit_config.patch_size
            self.vit_max_num_patch_per_side = config.vit_max_num_patch_per_side
            self.vit_hidden_size = config.vit_config.hidden_size
            self.connector = MLPconnector(self.vit_hidden_size, self.hidden_size, config.connector_act)
            self.vit_pos_embed = PositionEmbedding(self.vit_max_num_patch_per_side, self.hidden_size, self.device, sdt)
        self.get_flattened_position_ids = (get_flattened_position_ids_interpolate if config.interpolate_pos
                                           else get_flattened_position_ids_extrapolate)
        self.use_cuda_graph = True
        self._graph_cache: Dict[Any, Any] = {}

    def eval(self):
        return self

    # ------------------------------------------------------------------------------------------
    # weights (reference key schema)
    # ------------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = False):
        """Reference key schema. strict=True: every expected tensor must be present and nothing may be left over
        (KeyError lists both); strict=False: unexpected keys are ignored, but the heads this configuration needs
        (vae2llm / llm2vae / time_embedder for visual_gen, connector for visual_und) must still be there — a model with
        `None` weights would only fail later inside a kernel call."""
        lm_sd = {k[len("language_model."):]: v for k, v in sd.items() if k.startswith("language_model.")}
        unexpected = ["language_model." + k for k in self.language_model.load_state_dict(lm_sd, strict=False)]
        used = set()
        missing: List[str] = []

        def need(keys):
            miss = [k for k in keys if k not in sd]
            missing.extend(miss)
            used.update(keys)
            return not miss

        if self.config.visual_gen:
            if need(["time_embedder.mlp.0.weight", "time_embedder.mlp.0.bias", "time_embedder.mlp.2.weight",
                     "time_embedder.mlp.2.bias"]):
                self.time_embedder.load(sd, "time_embedder.", self.device)
            if need(["vae2llm.weight", "vae2llm.bias"]):
                self.vae2llm.load(sd, "vae2llm.", self.device)
            if need(["llm2vae.weight", "llm2vae.bias"]):
                self.llm2vae.load(sd, "llm2vae.", self.device)
            self.latent_pos_embed.load(sd.get("latent_pos_embed.pos_embed"))   # frozen sincos table: optional
            used.add("latent_pos_embed.pos_embed")
        if self.config.visual_und:
            if need(["connector.fc1.weight", "connector.fc1.bias", "connector.fc2.weight", "connector.fc2.bias"]):
                self.connector.load(sd, "connector.", self.device)
            self.vit_pos_embed.load(sd.get("vit_pos_embed.pos_embed"))
            used.add("vit_pos_embed.pos_embed")
            vit_sd = {k[len("vit_model."):]: v for k, v in sd.items() if k.startswith("vit_model.")}
            used.update("vit_model." + k for k in vit_sd)
            if self.vit_model is not None and hasattr(self.vit_model, "load_state_dict"):
                if vit_sd:
                    self.vit_model.load_state_dict(vit_sd)
                else:
                    missing.append("vit_model.*")
        unexpected += [k for k in sd if not k.startswith("language_model.") and k not in used]
        if missing:
            raise KeyError(f"bagel_b200.Bagel.load_state_dict: missing weights {missing[:12]}"
                           + (f" (+{len(missing) - 12} more)" if len(missing) > 12 else ""))
        if strict and unexpected:
            raise KeyError(f"bagel_b200.Bagel.load_state_dict: unexpected keys {unexpected[:12]}")
        return self

    # ------------------------------------------------------------------------------------------
    # packers
    # ------------------------------------------------------------------------------------------
    def prepare_prompts(self, curr_kvlens, curr_rope, prompts, tokenizer, new_token_ids):
        bos, eos = new_token_ids["bos_token_id"], new_token_ids["eos_token_id"]
        ids = [[bos] + list(tokenizer.encode(p)) + [eos] for p in prompts]
        tl = torch.tensor([len(x) for x in ids], dtype=torch.int64)
        cl = torch.tensor(list(curr_kvlens), dtype=torch.int64)
        rope = torch.tensor(list(curr_rope), dtype=torch.int64)
        sample_start = torch.cumsum(cl + tl, 0) - (cl + tl)
        generation_input = {
            "text_token_lens": tl.to(torch.int),
            "packed_text_ids": torch.tensor([t for x in ids for t in x], dtype=torch.long),
            "packed_text_position_ids": _ranges(rope, tl),
            "packed_text_indexes": _ranges(sample_start + cl, tl),
            "packed_key_value_indexes": _ranges(sample_start, cl),
            "key_values_lens": cl.to(torch.int),
        }
        return generation_input, (cl + tl).tolist(), (rope + tl).tolist()

    def prepare_vae_latent(self, curr_kvlens, curr_rope, image_sizes, new_token_ids):
        ds = self.latent_downsample
        cl = torch.tensor(list(curr_kvlens), dtype=torch.int64)
        rope = torch.tensor(list(curr_rope), dtype=torch.int64)
        ntok = torch.tensor([(H // ds) * (W // ds) for H, W in image_sizes], dtype=torch.int64)
        ql = ntok + 2
        q_start = torch.cumsum(ql, 0) - ql
        b_start = torch.cumsum(cl + ql, 0) - (cl + ql)
        B = len(image_sizes)
        dim = self.latent_channel * self.latent_patch_size ** 2
        # init noise: drawn per sample, in sample order, from the global CPU generator exactly like the reference
        noises = [torch.randn(int(n), dim) for n in ntok]
        pos = [self.get_flattened_position_ids(H, W, ds, max_num_patches_per_side=self.max_latent_size)
               for H, W in image_sizes]
        generation_input = {
            "packed_text_ids": torch.tensor([new_token_ids["start_of_image"], new_token_ids["end_of_image"]] * B,
                                            dtype=torch.long),
            "packed_text_indexes": torch.stack([q_start, q_start + ntok + 1], dim=1).reshape(-1),
            "packed_init_noises": torch.cat(noises, dim=0),
            "packed_vae_position_ids": torch.cat(pos, dim=0),
            "packed_vae_token_indexes": _ranges(q_start + 1, ntok),
            "packed_seqlens": ql.to(torch.int),
            "packed_position_ids": torch.repeat_interleave(rope, ql),
            "key_values_lens": cl.to(torch.int),
            "packed_indexes": _ranges(b_start + cl, ql),
            "packed_key_value_indexes": _ranges(b_start, cl),
        }
        return generation_input

    def prepare_vae_latent_cfg(self, curr_kvlens, curr_rope, image_sizes):
        ds = self.latent_downsample
        cl = torch.tensor(list(curr_kvlens), dtype=torch.int64)
        rope = torch.tensor(list(curr_rope), dtype=torch.int64)
        ql = torch.tensor([(H // ds) * (W // ds) + 2 for H, W in image_sizes], dtype=torch.int64)
        b_start = torch.cumsum(cl + ql, 0) - (cl + ql)
        return {
            "cfg_packed_position_ids": torch.repeat_interleave(rope, ql),
            "cfg_key_values_lens": cl.to(torch.int),
            "cfg_packed_query_indexes": _ranges(b_start + cl, ql),
            "cfg_packed_key_value_indexes": _ranges(b_start, cl),
        }

    def prepare_start_tokens(self, curr_kvlens, curr_rope, new_token_ids):
        cl = torch.tensor(list(curr_kvlens), dtype=torch.int64)
        # NB: like the reference (:909-927) these indexes do NOT leave a slot for the query token; generate_text
        # shifts sample i by i at the first step (:955-958).
        b_start = torch.cumsum(cl, 0) - cl
        B = len(curr_kvlens)
        return {
            "packed_start_tokens": torch.tensor([new_token_ids["bos_token_id"]] * B, dtype=torch.long),
            "packed_query_position_ids": torch.tensor(list(curr_rope), dtype=torch.long),
            "key_values_lens": cl.to(torch.int),
            "packed_key_value_indexes": _ranges(b_start, cl),
        }

    # ------------------------------------------------------------------------------------------
    # prefill
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_cache_update_text(self, past_key_values: NaiveCache, packed_text_ids, packed_text_position_ids,
                                  text_token_lens, packed_text_indexes, packed_key_value_indexes, key_values_lens):
        emb = self.language_model.model.embed_tokens(packed_text_ids)
        out = self.language_model.forward_inference(
            packed_query_sequence=emb, query_lens=text_token_lens,
            packed_query_position_ids=packed_text_position_ids, packed_query_indexes=packed_text_indexes,
            past_key_values=past_key_values, packed_key_value_indexes=packed_key_value_indexes,
            key_values_lens=key_values_lens, update_past_key_values=True, is_causal=True, mode="und")
        return out.past_key_values

    # ------------------------------------------------------------------------------------------
    # image understanding context: SigLIP tokens (reference bagel.py:299-415)
    # ------------------------------------------------------------------------------------------
    def prepare_vit_images(self, curr_kvlens, curr_rope, images, transforms, new_token_ids):
        cl = torch.tensor(list(curr_kvlens), dtype=torch.int64)
        rope = torch.tensor(list(curr_rope), dtype=torch.int64)
        tokens, pos = [], []
        for image in images:
            if hasattr(transforms, "patches") and hasattr(image, "size") and not torch.is_tensor(image):
                # device-side image path (transforms.DeviceImageTransform): resize + normalise + patchify on the GPU,
                # bit-identical to the host path below
                w_, h_ = transforms.resize_transform.target_size(*image.size)
                tokens.append(transforms.patches(image, self.vit_patch_size))
            else:
                t = transforms(image)
                h_, w_ = t.size(1), t.size(2)
                tokens.append(patchify(t, self.vit_patch_size))
            pos.append(self.get_flattened_position_ids(h_, w_, self.vit_patch_size,
                                                       max_num_patches_per_side=self.vit_max_num_patch_per_side))
        ntok = torch.tensor([x.shape[0] for x in tokens], dtype=torch.int64)
        ql = ntok + 2
        q_start = torch.cumsum(ql, 0) - ql
        b_start = torch.cumsum(cl + ql, 0) - (cl + ql)
        B = len(images)
        generation_input = {
            "packed_text_ids": torch.tensor([new_token_ids["start_of_image"], new_token_ids["end_of_image"]] * B,
                                            dtype=torch.long),
            "packed_text_indexes": torch.stack([q_start, q_start + ntok + 1], dim=1).reshape(-1),
            "vit_token_seqlens": ntok.to(torch.int),
            "packed_vit_tokens": torch.cat(tokens, dim=0),
            "packed_vit_position_ids": torch.cat(pos, dim=0),
            "packed_vit_token_indexes": _ranges(q_start + 1, ntok),
            "packed_position_ids": torch.repeat_interleave(rope, ql),
            "packed_seqlens": ql.to(torch.int),
            "packed_indexes": _ranges(b_start + cl, ql),
            "packed_key_value_indexes": _ranges(b_start, cl),
            "key_values_lens": cl.to(torch.int),
        }
        # an image block shares ONE rope position; the counter then advances by 1 (reference :340-343)
        return generation_input, (cl + ql).tolist(), (rope + 1).tolist()

    @torch.no_grad()
    def forward_cache_update_vit(self, past_key_values: NaiveCache, packed_text_ids, packed_text_indexes,
                                 packed_vit_tokens, packed_vit_token_indexes, packed_vit_position_ids,
                                 vit_token_seqlens, packed_position_ids, packed_seqlens, packed_indexes,
                                 packed_key_value_indexes, key_values_lens):
        dev = self.device
        lm = self.language_model.model
        if self.dtype_mode == "B":
            raise NotImplementedError("dtype_mode='B' covers the LM, the text / VAE prefills and the sampler; the SigLIP tower "
                                      "runs its bf16-stream path only (mode A)")
        n = int(torch.as_tensor(packed_seqlens).sum())
        seq = torch.zeros((n, self.hidden_size), dtype=BF16, device=dev)
        emb = lm.embed_tokens(torch.as_tensor(packed_text_ids))
        ops.copy_rows(emb, seq, dst_rows=torch.as_tensor(packed_text_indexes).to(dev, torch.int32))
        vl = torch.as_tensor(vit_token_seqlens).to("cpu", torch.int64)
        cu = torch.cat([torch.zeros(1, dtype=torch.int64), vl.cumsum(0)]).to(torch.int32)
        feats = self.vit_model(packed_pixel_values=packed_vit_tokens,
                               packed_flattened_position_ids=packed_vit_position_ids, cu_seqlens=cu,
                               max_seqlen=int(vl.max()))
        feats = self.connector(feats)
        # + vit_pos_embed[pos], scattered to the image rows of the packed sequence (reference :390-395)
        ops.latent_embed_add(feats, None, self.vit_pos_embed.pos_embed,
                             torch.as_tensor(packed_vit_position_ids).to(dev, torch.int64).contiguous(), seq,
                             torch.as_tensor(packed_vit_token_indexes).to(dev, torch.int32))
        out = self.language_model.forward_inference(
            packed_query_sequence=seq, query_lens=packed_seqlens, packed_query_position_ids=packed_position_ids,
            packed_query_indexes=packed_indexes, past_key_values=past_key_values,
            packed_key_value_indexes=packed_key_value_indexes, key_values_lens=key_values_lens,
            update_past_key_values=True, is_causal=False, mode="und")
        return out.past_key_values

    # ------------------------------------------------------------------------------------------
    # image editing context: clean VAE latents at t=0 (reference bagel.py:417-550)
    # ------------------------------------------------------------------------------------------
    def prepare_vae_images(self, curr_kvlens, curr_rope, images, transforms, new_token_ids, timestep=0):
        ds = self.latent_downsample
        cl = torch.tensor(list(curr_kvlens), dtype=torch.int64)
        rope = torch.tensor(list(curr_rope), dtype=torch.int64)
        tensors = [transforms(im) for im in images]
        shapes = [(t.shape[1] // ds, t.shape[2] // ds) for t in tensors]
        ntok = torch.tensor([h * w for h, w in shapes], dtype=torch.int64)
        ql = ntok + 2
        q_start = torch.cumsum(ql, 0) - ql
        b_start = torch.cumsum(cl + ql, 0) - (cl + ql)
        B = len(images)
        C = tensors[0].shape[0]
        Hm, Wm = max(t.shape[1] for t in tensors), max(t.shape[2] for t in tensors)
        padded = torch.zeros((B, C, Hm, Wm), device=tensors[0].device)   # stays on the GPU with DeviceImageTransform
        for i, t in enumerate(tensors):
            padded[i, :, : t.shape[1], : t.shape[2]] = t
        generation_input = {
            "padded_images": padded,
            "patchified_vae_latent_shapes": shapes,
            "packed_vae_position_ids": torch.cat([self.get_flattened_position_ids(
                t.size(1), t.size(2), ds, max_num_patches_per_side=self.max_latent_size) for t in tensors], dim=0),
            "packed_timesteps": torch.tensor([timestep]),
            "packed_vae_token_indexes": _ranges(q_start + 1, ntok),
            "packed_text_ids": torch.tensor([new_token_ids["start_of_image"], new_token_ids["end_of_image"]] * B,
                                            dtype=torch.long),
            "packed_text_indexes": torch.stack([q_start, q_start + ntok + 1], dim=1).reshape(-1),
            "packed_position_ids": torch.repeat_interleave(rope, ql),
            "packed_seqlens": ql.to(torch.int),
            "packed_indexes": _ranges(b_start + cl, ql),
            "packed_key_value_indexes": _ranges(b_start, cl),
            "key_values_lens": cl.to(torch.int),
        }
        return generation_input, (cl + ql).tolist(), (rope + 1).tolist()

    @torch.no_grad()
    def forward_cache_update_vae(self, vae_model, past_key_values: NaiveCache, padded_images,
                                 patchified_vae_latent_shapes, packed_vae_position_ids, packed_timesteps,
                                 packed_vae_token_indexes, packed_text_ids, packed_text_indexes, packed_position_ids,
                                 packed_seqlens, packed_indexes, key_values_lens, packed_key_value_indexes):
        dev = self.device
        lm = self.language_model.model
        modeB = self.dtype_mode == "B"
        n = int(torch.as_tensor(packed_seqlens).sum())
        seq = torch.zeros((n, self.hidden_size), dtype=lm.stream_dtype, device=dev)
        emb = lm.embed_tokens(torch.as_tensor(packed_text_ids))
        (ops.copy_rows_f32 if modeB else ops.copy_rows)(emb, seq, dst_rows=torch.as_tensor(packed_text_indexes).to(dev, torch.int32))
        latents = vae_model.encode(padded_images)                      # [B, z, Hm/8, Wm/8]
        p, zc = self.latent_patch_size, self.latent_channel
        rows = []
        for lat, (h, w) in zip(latents, patchified_vae_latent_shapes):  # 2x2 patchify, (p, q, c) order (:517-518)
            lat = lat[:, : h * p, : w * p].reshape(zc, h, p, w, p)
            rows.append(lat.permute(1, 3, 2, 4, 0).reshape(h * w, p * p * zc))
        packed_latent = torch.cat(rows, dim=0).to(dev, BF16).contiguous()
        proj = ops.gemm(packed_latent, self.vae2llm.weight, bias=self.vae2llm.bias)
        t_emb = self.time_embedder(torch.as_tensor(packed_timesteps).to(dev, torch.float32).reshape(-1)[:1])
        (ops.latent_embed_add_f32 if modeB else ops.latent_embed_add)(
            proj, t_emb[0], self.latent_pos_embed.pos_embed,
            torch.as_tensor(packed_vae_position_ids).to(dev, torch.int64).contiguous(), seq,
            torch.as_tensor(packed_vae_token_indexes).to(dev, torch.int32))
        extra = {}
        if self.use_moe:
            extra = dict(mode="gen", packed_vae_token_indexes=packed_vae_token_indexes,
                         packed_text_indexes=packed_text_indexes)
        out = self.language_model.forward_inference(
            packed_query_sequence=seq, query_lens=packed_seqlens, packed_query_position_ids=packed_position_ids,
            packed_query_indexes=packed_indexes, past_key_values=past_key_values, key_values_lens=key_values_lens,
            packed_key_value_indexes=packed_key_value_indexes, update_past_key_values=True, is_causal=False, **extra)
        return out.past_key_values

    # ------------------------------------------------------------------------------------------
    # rectified-flow sampler
    # ------------------------------------------------------------------------------------------
    def _build_flow_plan(self, branches: List[Dict[str, Any]], packed_seqlens, packed_vae_token_indexes,
                         packed_text_indexes):
        """One ForwardPlan covering all CFG branches as extra samples of the packed batch, plus merged KV
        buffers with every branch's (read-only) context rows placed once."""
        lm = self.language_model.model
        n = int(torch.as_tensor(packed_seqlens).sum())
        ql, pos, qidx, kvl, kvidx, vae, txt = [], [], [], [], [], [], []
        row_off = 0
        ctx_pairs = []
        for b, br in enumerate(branches):
            kl = torch.as_tensor(br["key_values_lens"]).to("cpu", torch.int64)
            has_ctx = br["past_key_values"] is not None and br["past_key_values"].key_cache[0] is not None
            if not has_ctx:
                kl = torch.zeros_like(kl)
            total_b = n + int(kl.sum())
            ql.append(torch.as_tensor(packed_seqlens).to("cpu", torch.int64))
            pos.append(torch.as_tensor(br["packed_position_ids"]).to("cpu", torch.int64))
            qidx.append(torch.as_tensor(br["packed_query_indexes"]).to("cpu", torch.int64) + row_off)
            kvl.append(kl)
            ki = torch.as_tensor(br["packed_key_value_indexes"]).to("cpu", torch.int64) if has_ctx else torch.zeros(0, dtype=torch.int64)
            kvidx.append(ki + row_off)
            if has_ctx:
                ctx_pairs.append((br["past_key_values"], (ki + row_off).to(self.device, torch.int32)))
            vae.append(torch.as_tensor(packed_vae_token_indexes).to("cpu", torch.int64) + b * n)
            txt.append(torch.as_tensor(packed_text_indexes).to("cpu", torch.int64) + b * n)
            row_off += total_b
        plan = lm.make_plan(query_lens=torch.cat(ql), position_ids=torch.cat(pos),
                            packed_query_indexes=torch.cat(qidx), key_values_lens=torch.cat(kvl),
                            packed_key_value_indexes=torch.cat(kvidx), is_causal=False,
                            mode="gen" if self.use_moe else "und",
                            packed_vae_token_indexes=torch.cat(vae), packed_text_indexes=torch.cat(txt))
        kbuf, vbuf = lm.alloc_kv(plan)
        cfg = lm.config
        w = cfg.num_key_value_heads * cfg.head_dim
        for cache, rows in ctx_pairs:
            m = rows.numel()
            for li in range(cfg.num_hidden_layers):
                ops.copy_rows(cache.key_cache[li].reshape(m, w), kbuf[li], dst_rows=rows, M=m)
                ops.copy_rows(cache.value_cache[li].reshape(m, w), vbuf[li], dst_rows=rows, M=m)
        return plan, kbuf, vbuf

    def _velocity(self, st: Dict[str, Any], key: str, t_row: torch.Tensor, x_src: torch.Tensor, head: bool = True) -> int:
        """Latent-in (bagel.py:796-806) -> packed LM call over all CFG branches -> llm2vae (:832). Fills
        st['v_all'][b*n:(b+1)*n] with branch b's outputs for every packed row; returns the branch count.
        head=False stops after the last decoder layer (TaylorSeer caches / replaces that tensor)."""
        lm = self.language_model.model
        plan, kbuf, vbuf, nb = st[key]
        n = st["n"]
        modeB = self.dtype_mode == "B"
        embed_add = ops.latent_embed_add_f32 if modeB else ops.latent_embed_add
        copy_rows = ops.copy_rows_f32 if modeB else ops.copy_rows
        seq = lm._buf("xa", plan.n, self.hidden_size, lm.stream_dtype)
        ops.cast_f32_to_bf16(x_src, out=st["x_bf16"])
        ops.gemm(st["x_bf16"], self.vae2llm.weight, bias=self.vae2llm.bias, out=st["proj"])
        for b in range(nb):
            embed_add(st["proj"], t_row, self.latent_pos_embed.pos_embed, st["vae_pos"], seq[b * n:(b + 1) * n], st["vae_rows"])
            copy_rows(st["text_emb"], seq[b * n:(b + 1) * n], dst_rows=st["text_rows"])
        lm.run_layers(seq, plan, kbuf, vbuf, final_norm=False)
        if head:
            self._velocity_head(st, key)
        return nb

    def _velocity_head(self, st: Dict[str, Any], key: str):
        """Final norm + llm2vae (bagel.py:832) over the hidden state in the LM's "xa" workspace."""
        lm = self.language_model.model
        plan, _, _, nb = st[key]
        out = lm.final_norm(plan, for_linear=True)
        ops.gemm(out, self.llm2vae.weight, bias=self.llm2vae.bias, out=st["v_all"][: nb * st["n"]])

    def _cfg_update(self, st: Dict[str, Any], nb: int, scales: Tuple[float, float], renorm_min: float,
                    renorm_type: str, x_dst: torch.Tensor, dt: float, dt_dev: Optional[torch.Tensor] = None):
        """CFG combine + renorm (bagel.py:873-905) fused with the Euler update x -= v*dt (:746)."""
        n, v = st["n"], st["v_all"]
        sT, sI = scales
        v_text = v[n:2 * n] if (nb >= 2 and sT > 1.0) else None
        v_img = v[2 * n:3 * n] if (nb >= 3 and sI > 1.0 and v_text is not None) else None
        ops.cfg_euler_step(v[:n], v_text, v_img, st["vae_rows"], x_dst, st["norms"],
                           sT if v_text is not None else 1.0, sI if v_img is not None else 1.0, renorm_min,
                           renorm_type, dt, dt_dev)

    def _flow_state(self, x_t, packed_seqlens, packed_vae_token_indexes, packed_text_indexes,
                    packed_vae_position_ids, packed_text_ids, nb: int) -> Dict[str, Any]:
        dev = self.device
        lm = self.language_model.model
        n = int(torch.as_tensor(packed_seqlens).sum())
        vae_idx = torch.as_tensor(packed_vae_token_indexes).to("cpu", torch.int64)
        txt_idx = torch.as_tensor(packed_text_indexes).to("cpu", torch.int64)
        M = int(vae_idx.numel())
        st: Dict[str, Any] = {"n": n, "M": M, "vae_idx": vae_idx, "txt_idx": txt_idx}
        st["x"] = x_t.to(dev, torch.float32, non_blocking=True).contiguous().clone()
        st["x_bf16"] = torch.empty((M, self.patch_latent_dim), dtype=BF16, device=dev)
        st["proj"] = torch.empty((M, self.hidden_size), dtype=BF16, device=dev)
        st["v_all"] = torch.empty((nb * n, self.patch_latent_dim), dtype=BF16, device=dev)
        st["norms"] = torch.zeros(2, dtype=torch.float32, device=dev)
        st["vae_rows"] = vae_idx.to(dev, torch.int32)
        st["text_rows"] = txt_idx.to(dev, torch.int32)
        st["vae_pos"] = torch.as_tensor(packed_vae_position_ids).to(dev, torch.int64).contiguous()
        st["text_emb"] = lm.embed_tokens(torch.as_tensor(packed_text_ids))
        return st

    @torch.no_grad()
    def make_flow_runner(self, packed_text_ids, packed_text_indexes, packed_init_noises, packed_vae_position_ids,
                         packed_vae_token_indexes, packed_seqlens, packed_position_ids, packed_indexes,
                         past_key_values: NaiveCache, key_values_lens, packed_key_value_indexes,
                         num_timesteps: int = 24, timestep_shift: float = 1.0, cfg_renorm_min: float = 0.0,
                         cfg_renorm_type: str = "global", cfg_interval: Optional[Sequence[float]] = (0, 1),
                         cfg_text_scale: float = 1.0, cfg_text_packed_query_indexes=None,
                         cfg_text_packed_position_ids=None, cfg_text_past_key_values: Optional[NaiveCache] = None,
                         cfg_text_key_values_lens=None, cfg_text_packed_key_value_indexes=None,
                         cfg_img_scale: float = 1.0, cfg_img_packed_query_indexes=None,
                         cfg_img_packed_position_ids=None, cfg_img_past_key_values: Optional[NaiveCache] = None,
                         cfg_img_key_values_lens=None, cfg_img_packed_key_value_indexes=None,
                         cfg_type: str = "parallel", enable_taylorseer: bool = False) -> "FlowRunner":
        """Plan a whole denoising run (same arguments as generate_image); FlowRunner.step(i) then executes
        velocity evaluation + CFG + Euler update number i as a sync-free kernel sequence."""
        if cfg_renorm_type not in ops.RENORM:
            raise NotImplementedError(f"{cfg_renorm_type} is not supported")
        if enable_taylorseer and self.dtype_mode == "B":
            raise NotImplementedError("enable_taylorseer=True is implemented for dtype_mode='A' only (bf16 factor planes)")
        dev = self.device

        # ---- schedule (host; identical arithmetic to the reference :693-696) ----
        ts = torch.linspace(1, 0, num_timesteps)
        ts = timestep_shift * ts / (1 + (timestep_shift - 1) * ts)
        dts = ts[:-1] - ts[1:]
        ts = ts[:-1]
        cfg_on = [bool(t > cfg_interval[0] and t <= cfg_interval[1]) for t in ts]

        main = dict(packed_position_ids=packed_position_ids, packed_query_indexes=packed_indexes,
                    past_key_values=past_key_values, key_values_lens=key_values_lens,
                    packed_key_value_indexes=packed_key_value_indexes)
        branches = [main]
        if cfg_text_scale > 1.0:
            branches.append(dict(packed_position_ids=cfg_text_packed_position_ids,
                                 packed_query_indexes=cfg_text_packed_query_indexes,
                                 past_key_values=cfg_text_past_key_values, key_values_lens=cfg_text_key_values_lens,
                                 packed_key_value_indexes=cfg_text_packed_key_value_indexes))
            if cfg_img_scale > 1.0:  # the reference consumes the image branch only inside the text-CFG block (:873)
                branches.append(dict(packed_position_ids=cfg_img_packed_position_ids,
                                     packed_query_indexes=cfg_img_packed_query_indexes,
                                     past_key_values=cfg_img_past_key_values, key_values_lens=cfg_img_key_values_lens,
                                     packed_key_value_indexes=cfg_img_packed_key_value_indexes))
        nbmax = len(branches)
        # every LM workspace at its final size BEFORE the first launch / graph capture: the 'full' (all-branch) steps
        # need nbmax*n rows, the 'main' steps n — growing a buffer in between would free memory a captured graph replays into
        n_rows = int(torch.as_tensor(packed_seqlens).sum())
        self.language_model.model.reserve(nbmax * n_rows, nbmax * int(torch.as_tensor(packed_text_indexes).numel()))
        st = self._flow_state(packed_init_noises, packed_seqlens, packed_vae_token_indexes, packed_text_indexes,
                              packed_vae_position_ids, packed_text_ids, nbmax)
        st["t_emb"] = self.time_embedder(ts.to(dev))  # every timestep of the run at once: [num_timesteps-1, H]
        if any(cfg_on) and nbmax > 1:
            st["full"] = (*self._build_flow_plan(branches, packed_seqlens, st["vae_idx"], st["txt_idx"]), nbmax)
        if (not all(cfg_on)) or nbmax == 1:
            st["main"] = (*self._build_flow_plan(branches[:1], packed_seqlens, st["vae_idx"], st["txt_idx"]), 1)
        return FlowRunner(self, st, dts.tolist(), cfg_on, (cfg_text_scale, cfg_img_scale), cfg_renorm_min,
                          cfg_renorm_type, nbmax, torch.as_tensor(packed_seqlens).to("cpu", torch.int64),
                          enable_taylorseer=enable_taylorseer)

    @torch.no_grad()
    def generate_image(self, *args, **kwargs):
        """Rectified-flow Euler sampler (reference bagel.py:644-754), same signature as the reference; returns the
        tuple of per-sample latents [h*w, 64] fp32 (on the model's device)."""
        runner = self.make_flow_runner(*args, **kwargs)
        for i in range(runner.num_steps):
            runner.step(i)
        return runner.latents()

    # ------------------------------------------------------------------------------------------
    # text decode
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate_text(self, past_key_values: NaiveCache, packed_key_value_indexes, key_values_lens,
                      packed_start_tokens, packed_query_position_ids, max_length: int, do_sample: bool = False,
                      temperature: float = 1.0, end_token_id: Optional[int] = None):
        """Greedy / sampled decode, one token per sample per step (reference bagel.py:930-1000). Returns [steps, B]
        token ids on the model's device. As in the reference, generation stops when SAMPLE 0 emits `end_token_id`
        (:996) and the stopping token is not returned.

        B200-first execution (the reference re-allocates and re-scatters the whole KV cache per layer per token and
        rebuilds index tensors with host loops): the KV cache is copied ONCE into per-sample slabs with room for
        `max_length` new tokens; sequence lengths, RoPE positions, write slots and the token history live on the
        device; a step is embedding gather -> 28 layers (fused QKV epilogue appends K/V in place, attention reads
        `seqused_k`) -> final norm -> lm_head -> argmax, captured once as a CUDA graph and replayed per token.
        The only host<->device traffic per step is the 8-byte EOS check the reference also performs."""
        dev = self.device
        lm = self.language_model.model
        if self.dtype_mode == "B":
            raise NotImplementedError("generate_text: the device-resident decode loop is implemented for dtype_mode='A'")
        cfg = lm.config
        L, H, Hq, Hk, D = cfg.num_hidden_layers, cfg.hidden_size, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
        w = Hk * D
        kv = torch.as_tensor(key_values_lens).to("cpu", torch.int64)
        B = int(kv.numel())
        if max_length <= 0:
            return torch.zeros((0, B), dtype=torch.int64, device=dev)
        cap = kv + max_length
        begin = torch.cumsum(cap, 0) - cap
        total = int(cap.sum())
        max_kv = int(cap.max())   # host upper bound of any sample's key count (sizes the key split of decode attention)
        has_ctx = past_key_values is not None and past_key_values.key_cache[0] is not None and int(kv.sum()) > 0
        # zero-filled, not torch.empty: attention multiplies the masked probabilities (exactly 0) with whatever sits in the
        # spare rows of a slab — 0 x NaN/Inf garbage would poison the output (the tcgen05 kernel fetches whole 128-key
        # blocks by TMA; only the single-query d=128 kernel clamps its loads to the rows in use)
        kbuf = torch.zeros((L, total, w), dtype=BF16, device=dev)
        vbuf = torch.zeros((L, total, w), dtype=BF16, device=dev)
        if has_ctx:
            n_ctx = int(kv.sum())
            dst = _ranges(begin, kv).to(dev, torch.int32)
            for li in range(L):
                ops.copy_rows(past_key_values.key_cache[li].reshape(n_ctx, w), kbuf[li], dst_rows=dst, M=n_ctx)
                ops.copy_rows(past_key_values.value_cache[li].reshape(n_ctx, w), vbuf[li], dst_rows=dst, M=n_ctx)
        k_begin = torch.cat([begin, torch.tensor([total])]).to(dev, torch.int32)
        cu_q = torch.arange(B + 1, dtype=torch.int32, device=dev)
        seq_len = kv.to(dev, torch.int32)
        pos = torch.as_tensor(packed_query_position_ids).to(dev, torch.int64).clone()
        tokens = torch.as_tensor(packed_start_tokens).to(dev, torch.int64).clone()
        tokens32 = tokens.to(torch.int32)
        history = torch.zeros((max_length, B), dtype=torch.int64, device=dev)
        step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        kv_rows = torch.zeros(B, dtype=torch.int32, device=dev)
        seqused = torch.zeros(B, dtype=torch.int32, device=dev)
        x = torch.empty((B, H), dtype=BF16, device=dev)
        logits = torch.empty((B, cfg.vocab_size), dtype=BF16, device=dev)
        eps = cfg.rms_norm_eps
        bufs = dict(xb=torch.empty_like(x), h=torch.empty_like(x),
                    qkv=torch.empty((B, (Hq + 2 * Hk) * D), dtype=BF16, device=dev),
                    q=torch.empty((B, Hq * D), dtype=BF16, device=dev), att=torch.empty((B, Hq * D), dtype=BF16, device=dev),
                    act=torch.empty((B, cfg.intermediate_size), dtype=BF16, device=dev), out=torch.empty_like(x))
        head = self.language_model.lm_head

        def body():
            """One decode step; every input/output is a fixed device buffer (graph-replayable)."""
            ops.copy_rows(lm.embed_tokens.weight, x, src_rows=tokens32, M=B)
            cos, sin = bufs.get("cos"), bufs.get("sin")
            ops.rope_table_into(pos, lm.inv_freq, cos, sin, True)
            ops.decode_prepare(k_begin, seq_len, kv_rows, seqused)
            xa, xb, h = x, bufs["xb"], bufs["h"]
            for li, layer in enumerate(lm.layers):
                e = layer.und
                ops.rmsnorm(xa, e.ln_in, None, None, eps, out=h)
                if lm.fused_qkv and D == 128 and B > 64:   # B <= 64: weight-streaming skinny GEMM + norm/RoPE kernel
                    ops.gemm_qkv_norm_rope(h, e.wqkv, e.bqkv, e.q_norm, e.k_norm, None, None, None, cos, sin, bufs["q"],
                                           kbuf[li], vbuf[li], kv_rows, Hq, Hk, eps, False)
                else:
                    ops.gemm(h, e.wqkv, bias=e.bqkv, out=bufs["qkv"])
                    ops.qk_norm_rope(bufs["qkv"], e.q_norm, e.k_norm, None, None, None, cos, sin, bufs["q"], kbuf[li],
                                     vbuf[li], kv_rows, Hq, Hk, D, eps, False)
                ops.attn_varlen(bufs["q"].view(B, Hq, D), kbuf[li].view(-1, Hk, D), vbuf[li].view(-1, Hk, D), cu_q, k_begin,
                                1, max_kv, True, out=bufs["att"].view(B, Hq, D), seqused_k=seqused)
                ops.gemm(bufs["att"], e.wo, resid=xa, epilogue=ops.EPI_RESID, out=xb)
                ops.rmsnorm(xb, e.ln_post, None, None, eps, out=h)
                ops.gemm(h, e.wgu, epilogue=ops.EPI_SWIGLU, out=bufs["act"])
                ops.gemm(bufs["act"], e.wd, resid=xb, epilogue=ops.EPI_RESID, out=xa)
            ops.rmsnorm(xa, lm.norm, None, None, eps, out=bufs["out"])
            ops.gemm(bufs["out"], head.weight, bias=head.bias, out=logits)

        def tail_greedy():
            ops.decode_advance(seq_len, pos, tokens, history, step_dev)   # history[step] = current tokens; lens += 1
            ops.argmax_rows(logits, tokens, tokens32)

        bufs["cos"] = torch.empty((B, D // 2), dtype=torch.float32, device=dev)
        bufs["sin"] = torch.empty((B, D // 2), dtype=torch.float32, device=dev)
        graph = None
        use_graph = bool(getattr(self, "use_cuda_graph", True)) and not do_sample
        steps = 0
        for step in range(max_length):
            if graph is not None:
                graph.replay()
            else:
                if use_graph and step == 1:
                    try:
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            body()
                            tail_greedy()
                        graph = g
                        graph.replay()
                    except Exception as e:
                        use_graph = False
                        import warnings
                        warnings.warn(f"bagel_b200: CUDA graph capture of the decode step failed, continuing eagerly: {e}")
                        torch.cuda.synchronize()
                        body()
                        tail_greedy()
                else:
                    body()
                    if do_sample:
                        ops.decode_advance(seq_len, pos, tokens, history, step_dev)
                        probs = torch.softmax(logits.float() / temperature, dim=-1)
                        nxt = torch.multinomial(probs, num_samples=1).squeeze(1)
                        tokens.copy_(nxt)
                        tokens32.copy_(nxt.to(torch.int32))
                    else:
                        tail_greedy()
            steps += 1
            if end_token_id is not None and int(tokens[0]) == end_token_id:
                break
        return history[:steps].clone()

    # ------------------------------------------------------------------------------------------
    # training-mode forward (losses only, no backward): reference Bagel.forward, bagel.py:101-229
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, sequence_length: int, packed_text_ids, packed_text_indexes, sample_lens: List[int],
                packed_position_ids, nested_attention_masks=None, split_lens: List[int] = None,
                attn_modes: List[str] = None, ce_loss_indexes=None, packed_label_ids=None, packed_vit_tokens=None,
                packed_vit_token_indexes=None, packed_vit_position_ids=None, vit_token_seqlens=None, padded_latent=None,
                patchified_vae_latent_shapes=None, packed_latent_position_ids=None, packed_vae_token_indexes=None,
                packed_timesteps=None, mse_loss_indexes=None, noise: Optional[torch.Tensor] = None):
        """The reference's training forward on a packed batch -> dict(mse=..., ce=...) (forward only: the B200 build has no
        autograd; useful for evaluation losses / distillation targets with the inference kernels).

        Attention structure (data/data_utils.py:13-40, 72-103): every sample is a list of splits that are 'causal' (text),
        'full' (an image everyone after it may look at) or 'noise' (a noised image only it itself sees); a split attends to
        all earlier non-noise splits of its sample plus itself. That is exactly a chain of prefill calls on a growing KV
        cache — so instead of a block-sparse mask over the whole packed sequence, each split runs through the varlen
        attention kernel against [cache || itself] (causal flag for text) and is appended to the cache unless it is
        noise. Experts as in training: text + ViT tokens -> und weights, VAE tokens -> gen weights, with the training
        modules' all-bf16 q/k-norm + RoPE flow (`train_numerics`). `split_lens` / `attn_modes` (the reference's
        flex-attention inputs) are required; `nested_attention_masks` carries no extra information and is ignored.
        `noise` overrides the torch.randn_like draw of :184."""
        if split_lens is None or attn_modes is None:
            raise NotImplementedError("Bagel.forward needs split_lens and attn_modes (dense nested_attention_masks alone "
                                      "cannot be mapped onto the varlen attention kernel)")
        if self.dtype_mode != "A":
            raise NotImplementedError("Bagel.forward is implemented for dtype_mode='A'")
        dev = self.device
        lm = self.language_model.model
        L = int(sequence_length)
        H = self.hidden_size
        i32 = lambda t: torch.as_tensor(t).to(dev, torch.int32).contiguous()
        seq = torch.zeros((L, H), dtype=BF16, device=dev)
        text_idx = torch.as_tensor(packed_text_indexes).to("cpu", torch.int64)
        ops.copy_rows(lm.embed_tokens(torch.as_tensor(packed_text_ids)), seq, dst_rows=i32(text_idx))
        kind = torch.zeros(L, dtype=torch.int8)          # 0 text (und), 1 ViT (und), 2 VAE (gen)
        if self.config.visual_und and packed_vit_tokens is not None:
            vl = torch.as_tensor(vit_token_seqlens).to("cpu", torch.int64)
            cu = torch.cat([torch.zeros(1, dtype=torch.int64), vl.cumsum(0)]).to(torch.int32)
            feats = self.connector(self.vit_model(packed_pixel_values=packed_vit_tokens,
                                                  packed_flattened_position_ids=packed_vit_position_ids, cu_seqlens=cu,
                                                  max_seqlen=int(vl.max())))
            ops.latent_embed_add(feats, None, self.vit_pos_embed.pos_embed,
                                 torch.as_tensor(packed_vit_position_ids).to(dev, torch.int64).contiguous(), seq,
                                 i32(packed_vit_token_indexes))
            kind[torch.as_tensor(packed_vit_token_indexes).to("cpu", torch.int64)] = 1
        mse = None
        if self.config.visual_gen and padded_latent is not None:
            p, zc = self.latent_patch_size, self.latent_channel
            rows = []
            for lat, (h, w) in zip(torch.as_tensor(padded_latent), patchified_vae_latent_shapes):
                lat = lat[:, : h * p, : w * p].reshape(zc, h, p, w, p)
                rows.append(lat.permute(1, 3, 2, 4, 0).reshape(h * w, p * p * zc))
            clean = torch.cat(rows, dim=0).to(dev, torch.float32)
            if noise is None:
                noise = torch.randn_like(clean)
            noise = noise.to(dev, torch.float32)
            t = torch.sigmoid(torch.as_tensor(packed_timesteps).to(dev, torch.float32))
            t = self.timestep_shift * t / (1 + (self.timestep_shift - 1) * t)
            x_t = ((1 - t[:, None]) * clean + t[:, None] * noise).contiguous()      # flow-matching interpolation (:185-187)
            proj = ops.gemm(ops.cast_f32_to_bf16(x_t), self.vae2llm.weight, bias=self.vae2llm.bias)
            t_emb = self.time_embedder(t)                                            # per token [M, H]
            vae_idx = torch.as_tensor(packed_vae_token_indexes).to("cpu", torch.int64)
            vae_rows = i32(vae_idx)
            vae_pos = torch.as_tensor(packed_latent_position_ids).to(dev, torch.int64).contiguous()
            off = 0
            for (h, w) in patchified_vae_latent_shapes:     # all tokens of an image share its timestep embedding row
                n = h * w
                ops.latent_embed_add(proj[off:off + n], t_emb[off], self.latent_pos_embed.pos_embed, vae_pos[off:off + n], seq,
                                     vae_rows[off:off + n])
                off += n
            kind[vae_idx] = 2
        # ---- the LM: one chained prefill per split ----
        hidden = torch.empty((L, H), dtype=BF16, device=dev)
        pos_all = torch.as_tensor(packed_position_ids).to("cpu", torch.int64)
        nl = lm.config.num_hidden_layers
        s_iter = iter(zip(split_lens, attn_modes))
        start = 0
        for slen in sample_lens:
            cache, kv_len, done = NaiveCache(nl), 0, 0
            while done < slen:
                n, amode = next(s_iter)
                if amode not in ("causal", "full", "noise"):
                    raise ValueError(f"unknown attn mode {amode!r}")
                r0 = start + done
                k = kind[r0:r0 + n]
                vae_rel = torch.nonzero(k == 2).reshape(-1)
                extra = {}
                if self.use_moe and vae_rel.numel():
                    extra = dict(mode="gen", packed_vae_token_indexes=vae_rel, packed_text_indexes=torch.nonzero(k != 2).reshape(-1))
                out = lm.forward_inference(
                    packed_query_sequence=seq[r0:r0 + n], query_lens=torch.tensor([n], dtype=torch.int32),
                    packed_query_position_ids=pos_all[r0:r0 + n], packed_query_indexes=torch.arange(kv_len, kv_len + n),
                    past_key_values=cache, key_values_lens=torch.tensor([kv_len], dtype=torch.int32),
                    packed_key_value_indexes=torch.arange(kv_len), update_past_key_values=(amode != "noise"),
                    is_causal=(amode == "causal"), train_numerics=True, **extra)
                hidden[r0:r0 + n] = out.packed_query_sequence
                if amode != "noise":
                    kv_len += n
                done += n
            if done != slen:
                raise ValueError("split_lens do not add up to sample_lens")
            start += slen
        self._last_hidden_state = hidden
        # ---- heads / losses (:214-227) ----
        if self.config.visual_gen and padded_latent is not None:
            mrows = torch.nonzero(torch.as_tensor(mse_loss_indexes).to("cpu")).reshape(-1)
            hm = torch.empty((mrows.numel(), H), dtype=BF16, device=dev)
            ops.copy_rows(hidden, hm, src_rows=i32(mrows))
            preds = ops.gemm(hm, self.llm2vae.weight, bias=self.llm2vae.bias)
            target = noise - clean                           # v_t = dx_t/dt = x_1 - x_0
            mse = (preds - target[t > 0]) ** 2
        ce = None
        if ce_loss_indexes is not None:
            crow = torch.nonzero(torch.as_tensor(ce_loss_indexes).to("cpu")).reshape(-1)
            hc = torch.empty((crow.numel(), H), dtype=BF16, device=dev)
            ops.copy_rows(hidden, hc, src_rows=i32(crow))
            logits = self.language_model.lm_head(hc)
            ce = torch.nn.functional.cross_entropy(logits.float(), torch.as_tensor(packed_label_ids).to(dev), reduction="none")
        return dict(mse=mse, ce=ce)

    __call__ = forward

    # ------------------------------------------------------------------------------------------
    # evaluation entry point: images + prompt -> text (reference bagel.py:1004-1075)
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def chat(self, tokenizer, new_token_ids, image_transform, images, prompt, max_length: int,
             do_sample: bool = False, temperature: float = 1.0):
        """Same call order as the reference: one SigLIP prefill per image (prepare_vit_images ->
        forward_cache_update_vit), then the prompt (prepare_prompts -> forward_cache_update_text), then greedy / sampled
        decode from <|im_start|> until <|im_end|>; returns the decoded answer between the two markers."""
        past_key_values = NaiveCache(self.config.llm_config.num_hidden_layers)
        newlens, new_rope = [0], [0]
        for image in images:
            generation_input, newlens, new_rope = self.prepare_vit_images(
                curr_kvlens=newlens, curr_rope=new_rope, images=[image], transforms=image_transform,
                new_token_ids=new_token_ids)
            past_key_values = self.forward_cache_update_vit(past_key_values, **generation_input)
        generation_input, newlens, new_rope = self.prepare_prompts(
            curr_kvlens=newlens, curr_rope=new_rope, prompts=[prompt], tokenizer=tokenizer, new_token_ids=new_token_ids)
        past_key_values = self.forward_cache_update_text(past_key_values, **generation_input)
        generation_input = self.prepare_start_tokens(newlens, new_rope, new_token_ids)
        unpacked_latent = self.generate_text(
            past_key_values=past_key_values, max_length=max_length, do_sample=do_sample, temperature=temperature,
            end_token_id=new_token_ids["eos_token_id"], **generation_input)
        output = tokenizer.decode(unpacked_latent[:, 0])
        return output.split("<|im_end|>")[0].split("<|im_start|>")[1]

    @torch.no_grad()
    def _forward_flow(self, x_t, timestep, packed_vae_token_indexes, packed_vae_position_ids, packed_text_ids,
                      packed_text_indexes, packed_indexes, packed_position_ids, packed_seqlens, key_values_lens,
                      past_key_values, packed_key_value_indexes, cfg_renorm_min=0.0, cfg_renorm_type="global",
                      cfg_text_scale=1.0, cfg_text_packed_position_ids=None, cfg_text_packed_query_indexes=None,
                      cfg_text_key_values_lens=None, cfg_text_past_key_values=None,
                      cfg_text_packed_key_value_indexes=None, cfg_img_scale=1.0, cfg_img_packed_position_ids=None,
                      cfg_img_packed_query_indexes=None, cfg_img_key_values_lens=None, cfg_img_past_key_values=None,
                      cfg_img_packed_key_value_indexes=None, cfg_type="parallel", model_pred_cache_dic=None,
                      model_pred_current=None, model_pred_text_cache_dic=None, model_pred_text_current=None,
                      model_pred_img_cache_dic=None, model_pred_img_current=None):
        """One velocity evaluation v_t [M, C] bf16 (reference :757-907). Implemented as a single Euler step of
        the fused path on x = 0 with dt = -1, which returns exactly the CFG-combined velocity."""
        if any(a is not None for a in (model_pred_cache_dic, model_pred_current, model_pred_text_cache_dic,
                                       model_pred_text_current, model_pred_img_cache_dic, model_pred_img_current)):
            # the reference threads its TaylorSeer caches through _forward_flow (:770-775); here the step cache belongs to
            # the planned run (FlowRunner) — a single stateless evaluation cannot honour it, so refuse instead of
            # silently computing a full step
            raise NotImplementedError("TaylorSeer caches are not accepted by _forward_flow; use "
                                      "generate_image(..., enable_taylorseer=True)")
        dev = self.device
        lm = self.language_model.model
        t = torch.as_tensor(timestep).to("cpu", torch.float32).reshape(-1)
        assert t.unique().numel() == 1
        main = dict(packed_position_ids=packed_position_ids, packed_query_indexes=packed_indexes,
                    past_key_values=past_key_values, key_values_lens=key_values_lens,
                    packed_key_value_indexes=packed_key_value_indexes)
        branches = [main]
        if cfg_text_scale > 1.0:
            branches.append(dict(packed_position_ids=cfg_text_packed_position_ids,
                                 packed_query_indexes=cfg_text_packed_query_indexes,
                                 past_key_values=cfg_text_past_key_values, key_values_lens=cfg_text_key_values_lens,
                                 packed_key_value_indexes=cfg_text_packed_key_value_indexes))
            if cfg_img_scale > 1.0:
                branches.append(dict(packed_position_ids=cfg_img_packed_position_ids,
                                     packed_query_indexes=cfg_img_packed_query_indexes,
                                     past_key_values=cfg_img_past_key_values, key_values_lens=cfg_img_key_values_lens,
                                     packed_key_value_indexes=cfg_img_packed_key_value_indexes))
        nb = len(branches)
        st = self._flow_state(x_t, packed_seqlens, packed_vae_token_indexes, packed_text_indexes,
                              packed_vae_position_ids, packed_text_ids, nb)
        st["t_emb"] = self.time_embedder(t[:1].to(dev))
        st["full"] = (*self._build_flow_plan(branches, packed_seqlens, st["vae_idx"], st["txt_idx"]), nb)
        self._velocity(st, "full", st["t_emb"][0], st["x"])
        v_out = torch.zeros_like(st["x"])  # x' = 0 - bf16(v * -1) = v
        self._cfg_update(st, nb, (cfg_text_scale, cfg_img_scale), cfg_renorm_min, cfg_renorm_type, v_out, -1.0)
        return v_out.to(BF16)


class FlowRunner:
    """A planned denoising run: x_t resident in HBM, one sync-free launch sequence per step. The per-step inputs
    that change (timestep embedding row, dt) live in fixed device buffers, so the launch sequence of a step is
    captured ONCE per branch set as a CUDA graph and replayed for the remaining steps."""

    def __init__(self, model: Bagel, st, dts, cfg_on, scales, renorm_min, renorm_type, nbmax, seqlens,
                 enable_taylorseer: bool = False):
        self.model, self.st, self.dts, self.cfg_on = model, st, dts, cfg_on
        self.scales, self.renorm_min, self.renorm_type, self.nbmax = scales, renorm_min, renorm_type, nbmax
        self.seqlens = seqlens
        self.num_steps = len(dts)
        dev = model.device
        self.dts_dev = torch.tensor(dts, dtype=torch.float32, device=dev)
        self.dt_cur = torch.zeros(1, dtype=torch.float32, device=dev)
        self.t_cur = torch.zeros_like(st["t_emb"][0])
        # CUDA graphs pay off only when a step is launch-bound. Capturing ~430 launches (twice, one graph per branch
        # set) costs ~0.7 s during which the GPU idles; at BAGEL-7B / 1024^2 / batch 8 a step is 0.8 s of GPU work
        # behind ~20 ms of asynchronous launches, so eager replay loses nothing there (measured: 39.4 s vs 40.1 s
        # per generate_image). Estimate the step from its linear-layer FLOPs at ~1 PFLOP/s.
        lcfg = model.language_model.model.config
        Hd, Id = lcfg.hidden_size, lcfg.intermediate_size
        qkv_o = (lcfg.num_attention_heads * 2 + lcfg.num_key_value_heads * 2) * lcfg.head_dim
        flops_step = 2.0 * nbmax * st["n"] * lcfg.num_hidden_layers * Hd * (qkv_o + 3 * Id)
        self.use_cuda_graph = bool(getattr(model, "use_cuda_graph", True)) and flops_step / 1.0e15 < 0.05
        self._graphs: Dict[str, Any] = {}
        self._graph_gen: Dict[str, int] = {}
        self._eager_done: Dict[str, int] = {}
        # TaylorSeer (reference bagel.py:680-684): one schedule per branch; factor planes of the last decoder
        # layer's output for every packed row of every branch, [7 orders, nbmax*n, H] bf16
        self.taylor = None
        if enable_taylorseer:
            from .taylorseer import TaylorSeerSchedule
            self.taylor = [TaylorSeerSchedule(len(dts) + 1) for _ in range(nbmax)]
            self.factors = torch.empty((TaylorSeerSchedule.MAX_ORDER + 1, nbmax * st["n"], model.hidden_size),
                                       dtype=BF16, device=dev)

    def _body(self, key: str):
        if self.taylor is not None:     # layers only; the cache update / extrapolation and the head follow eagerly
            self.model._velocity(self.st, key, self.t_cur, self.st["x"], head=False)
            return
        m, st = self.model, self.st
        on = key == "full"
        nb = m._velocity(st, key, self.t_cur, st["x"])
        m._cfg_update(st, nb, self.scales if on else (1.0, 1.0), self.renorm_min, self.renorm_type, st["x"], 0.0,
                      self.dt_cur)

    @torch.no_grad()
    def step(self, i: int):
        key = "full" if (self.cfg_on[i] and self.nbmax > 1) else "main"
        self.t_cur.copy_(self.st["t_emb"][i])
        self.dt_cur.copy_(self.dts_dev[i:i + 1])
        if self.taylor is not None:
            return self._step_taylorseer(key)
        self._launch(key)

    def _step_taylorseer(self, key: str):
        """One evaluation with the step cache. Branches are extra samples of one packed LM call, so the layers run
        whenever ANY active branch needs a fully computed step (with the reference's schedules the active branches
        always agree); branches on an extrapolated step get their rows of the last-layer output replaced."""
        m, st = self.model, self.st
        lm = m.language_model.model
        nb = st[key][3]
        n, H = st["n"], m.hidden_size
        scheds = self.taylor[:nb]
        types = [s.begin_step() for s in scheds]
        if any(t == "full" for t in types):
            self._launch(key)                                # embeddings + all decoder layers -> "xa"
        xa = lm._buf("xa", nb * n, H)
        for b, s in enumerate(scheds):
            rows = slice(b * n, (b + 1) * n)
            if s.type == "full":
                n_deriv, dist = s.full_update_args()
                ops.taylor_update(xa[rows], self.factors[:, rows], n_deriv, dist)
            else:
                n_f, x = s.taylor_args()
                ops.taylor_eval(self.factors[:, rows], n_f, x, xa[rows])
            s.end_step()
        m._velocity_head(st, key)
        on = key == "full"
        m._cfg_update(st, nb, self.scales if on else (1.0, 1.0), self.renorm_min, self.renorm_type, st["x"], 0.0,
                      self.dt_cur)

    def _launch(self, key: str):
        lm = self.model.language_model.model
        g = self._graphs.get(key)
        if g is not None:
            if self._graph_gen.get(key) == lm._ws_gen:
                g.replay()
                return
            # some LM workspace was re-allocated since the capture (another, larger forward ran between two steps):
            # the graph holds pointers into freed buffers -> drop it and capture again on the current ones
            del self._graphs[key]
        if self.use_cuda_graph and self._eager_done.get(key, 0) >= 1:
            try:  # everything is warm (workspaces allocated, kernel attributes set): capture this step
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self._body(key)
                self._graphs[key] = graph
                self._graph_gen[key] = lm._ws_gen
                graph.replay()          # capture does not execute the work
                return
            except Exception as e:  # capture is an optimisation; the eager launch sequence is the same work
                self.use_cuda_graph = False
                import warnings
                warnings.warn(f"bagel_b200: CUDA graph capture failed, continuing eagerly: {e}")
                torch.cuda.synchronize()
        self._body(key)
        self._eager_done[key] = self._eager_done.get(key, 0) + 1

    def latents(self):
        return self.st["x"].split((self.seqlens - 2).tolist())
