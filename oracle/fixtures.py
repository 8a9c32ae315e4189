"""TEST INFRASTRUCTURE — deterministic synthetic weights / inputs shared by the golden generator, the tests,
smoke() and bench.py. Nothing here reads /root/reference; tensors come from seeded CPU generators so the
same bits are produced in this container and on the GPU box (same torch build).

Random init follows the reference's recipe in spirit (normal weights, modeling_qwen2.py:563-572) but also
randomises norm weights and biases — with the stock init (ones / zeros) those code paths would be
untested — and gives llm2vae non-zero weights (the reference zero-inits it, bagel.py:96-99, which makes
v_t constant).
"""
from __future__ import annotations

from typing import Dict

import torch

from .qwen2_mot import LMConfig

TINY_LM = LMConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                   num_key_value_heads=2, vocab_size=1024)            # BASELINE.json configs[0]
TINY128_LM = LMConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                      num_key_value_heads=1, vocab_size=1024)         # same but head_dim 128 (7B's head_dim)
TINY_DENSE_LM = LMConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, vocab_size=1024, layer_module="Qwen2DecoderLayer")
TINY_MOE_LM = LMConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                       num_key_value_heads=1, vocab_size=1024, layer_module="Qwen2MoEDecoderLayer")   # head_dim 128
BAGEL_7B_LM = LMConfig(hidden_size=3584, intermediate_size=18944, num_hidden_layers=28, num_attention_heads=28,
                       num_key_value_heads=4, vocab_size=152064)      # BAGEL-7B-MoT llm_config.json


def _normal(gen, shape, std):
    return torch.randn(shape, generator=gen, dtype=torch.float32) * std


def lm_state_dict(cfg: LMConfig, seed: int = 0, dtype=torch.bfloat16, w_std: float = 0.05,
                  lm_head: bool = True) -> Dict[str, torch.Tensor]:
    """Reference key names (qwen2_navit.py Qwen2ForCausalLM state_dict) for cfg.layer_module: MoT layers carry every
    module twice ("" / "_moe_gen"), MoE layers only the MLP, dense layers nothing extra."""
    if cfg.layer_module != "Qwen2MoTDecoderLayer":
        sd = lm_state_dict(LMConfig(**{**cfg.__dict__, "layer_module": "Qwen2MoTDecoderLayer"}), seed, dtype, w_std, lm_head)
        moe = cfg.layer_module == "Qwen2MoEDecoderLayer"
        return {k: v for k, v in sd.items()
                if "_moe_gen" not in k or (moe and (".mlp_moe_gen." in k or k == "model.norm_moe_gen.weight"))}
    g = torch.Generator().manual_seed(seed)
    H, I, d = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
    Hq, Hk = cfg.num_attention_heads, cfg.num_key_value_heads
    sd = {"model.embed_tokens.weight": _normal(g, (cfg.vocab_size, H), 1.0)}
    for li in range(cfg.num_hidden_layers):
        p = f"model.layers.{li}."
        for sfx in ("", "_moe_gen"):
            a = p + "self_attn."
            sd[a + f"q_proj{sfx}.weight"] = _normal(g, (Hq * d, H), w_std)
            sd[a + f"q_proj{sfx}.bias"] = _normal(g, (Hq * d,), 0.1)
            sd[a + f"k_proj{sfx}.weight"] = _normal(g, (Hk * d, H), w_std)
            sd[a + f"k_proj{sfx}.bias"] = _normal(g, (Hk * d,), 0.1)
            sd[a + f"v_proj{sfx}.weight"] = _normal(g, (Hk * d, H), w_std)
            sd[a + f"v_proj{sfx}.bias"] = _normal(g, (Hk * d,), 0.1)
            sd[a + f"o_proj{sfx}.weight"] = _normal(g, (H, Hq * d), w_std)
            sd[a + f"q_norm{sfx}.weight"] = 1.0 + _normal(g, (d,), 0.1)
            sd[a + f"k_norm{sfx}.weight"] = 1.0 + _normal(g, (d,), 0.1)
            m = p + f"mlp{sfx}."
            sd[m + "gate_proj.weight"] = _normal(g, (I, H), w_std)
            sd[m + "up_proj.weight"] = _normal(g, (I, H), w_std)
            sd[m + "down_proj.weight"] = _normal(g, (H, I), w_std)
            sd[p + f"input_layernorm{sfx}.weight"] = 1.0 + _normal(g, (H,), 0.1)
            sd[p + f"post_attention_layernorm{sfx}.weight"] = 1.0 + _normal(g, (H,), 0.1)
    sd["model.norm.weight"] = 1.0 + _normal(g, (H,), 0.1)
    sd["model.norm_moe_gen.weight"] = 1.0 + _normal(g, (H,), 0.1)
    if lm_head:
        sd["lm_head.weight"] = _normal(g, (cfg.vocab_size, H), w_std)
    return {k: v.to(dtype) for k, v in sd.items()}


def bagel_extra_state_dict(hidden: int, patch_latent_dim: int = 64, seed: int = 1, dtype=torch.bfloat16,
                           w_std: float = 0.05) -> Dict[str, torch.Tensor]:
    """time_embedder / vae2llm / llm2vae parameters (bagel.py:78-82). The two frozen 2-D sincos tables
    (latent_pos_embed, vit_pos_embed) are deterministic and are built by the code under test."""
    g = torch.Generator().manual_seed(seed)
    sd = {
        "time_embedder.mlp.0.weight": _normal(g, (hidden, 256), w_std),
        "time_embedder.mlp.0.bias": _normal(g, (hidden,), 0.1),
        "time_embedder.mlp.2.weight": _normal(g, (hidden, hidden), w_std),
        "time_embedder.mlp.2.bias": _normal(g, (hidden,), 0.1),
        "vae2llm.weight": _normal(g, (hidden, patch_latent_dim), w_std * 2),
        "vae2llm.bias": _normal(g, (hidden,), 0.1),
        "llm2vae.weight": _normal(g, (patch_latent_dim, hidden), w_std),
        "llm2vae.bias": _normal(g, (patch_latent_dim,), 0.1),
    }
    return {k: v.to(dtype) for k, v in sd.items()}


def config1_inputs(cfg: LMConfig = TINY_LM, seq: int = 512, seed: int = 0, dtype=torch.bfloat16):
    """BASELINE.json configs[0] / SURVEY.md §8(d) cfg 1: one packed sequence of 512 rows."""
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.randn((seq, cfg.hidden_size), generator=g, dtype=torch.float32).to(dtype)
    return {
        "x": x,
        "query_lens": torch.tensor([seq], dtype=torch.int32),
        "und_position_ids": torch.arange(seq, dtype=torch.long),
        "gen_position_ids": torch.full((seq,), 7, dtype=torch.long),
        "query_indexes": torch.arange(seq, dtype=torch.long),
        "text_indexes": torch.tensor([0, seq - 1], dtype=torch.long),
        "vae_indexes": torch.arange(1, seq - 1, dtype=torch.long),
    }


def vit_state_dict(hidden: int, inter: int, layers: int, heads: int, llm_hidden: int, patch: int = 14,
                   max_side: int = 8, seed: int = 3, dtype=torch.bfloat16, w_std: float = 0.05) -> Dict[str, torch.Tensor]:
    """SigLIP NaViT tower (linear patch embedding, learned position table) + connector, reference key names."""
    g = torch.Generator().manual_seed(seed)
    pd = 3 * patch * patch
    p = "vit_model.vision_model."
    sd = {
        p + "embeddings.patch_embedding.weight": _normal(g, (hidden, pd), w_std),
        p + "embeddings.patch_embedding.bias": _normal(g, (hidden,), 0.1),
        p + "embeddings.position_embedding.weight": _normal(g, (max_side * max_side, hidden), 0.5),
        p + "post_layernorm.weight": 1.0 + _normal(g, (hidden,), 0.1),
        p + "post_layernorm.bias": _normal(g, (hidden,), 0.1),
    }
    for li in range(layers):
        q = p + f"encoder.layers.{li}."
        for n in ("q", "k", "v", "out"):
            sd[q + f"self_attn.{n}_proj.weight"] = _normal(g, (hidden, hidden), w_std * 2)
            sd[q + f"self_attn.{n}_proj.bias"] = _normal(g, (hidden,), 0.1)
        for n in ("layer_norm1", "layer_norm2"):
            sd[q + n + ".weight"] = 1.0 + _normal(g, (hidden,), 0.1)
            sd[q + n + ".bias"] = _normal(g, (hidden,), 0.1)
        sd[q + "mlp.fc1.weight"] = _normal(g, (inter, hidden), w_std * 2)
        sd[q + "mlp.fc1.bias"] = _normal(g, (inter,), 0.1)
        sd[q + "mlp.fc2.weight"] = _normal(g, (hidden, inter), w_std * 2)
        sd[q + "mlp.fc2.bias"] = _normal(g, (hidden,), 0.1)
    sd["connector.fc1.weight"] = _normal(g, (llm_hidden, hidden), w_std * 2)
    sd["connector.fc1.bias"] = _normal(g, (llm_hidden,), 0.1)
    sd["connector.fc2.weight"] = _normal(g, (llm_hidden, llm_hidden), w_std)
    sd["connector.fc2.bias"] = _normal(g, (llm_hidden,), 0.1)
    return {k: v.to(dtype) for k, v in sd.items()}


TINY_VIT = dict(hidden=144, inter=296, layers=2, heads=2)      # head_dim 72 like SigLIP-so400m (1152/16)


def vit_images(seed: int = 4):
    """Two 'already transformed' images [3,H,W] in [-1,1] with H,W multiples of the 14-px patch."""
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(3, 42, 56, generator=g) * 2 - 1, torch.rand(3, 28, 28, generator=g) * 2 - 1]


def vae_state_dict(ch: int = 128, ch_mult=(1, 2), num_res_blocks: int = 1, z_channels: int = 16, seed: int = 7,
                   dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random FLUX-VAE weights with the reference's key names (modeling/autoencoder.py module tree)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, co, ci, k):
        sd[name + ".weight"] = _normal(g, (co, ci, k, k), (1.0 / (ci * k * k)) ** 0.5)
        sd[name + ".bias"] = _normal(g, (co,), 0.05)

    def norm(name, c):
        sd[name + ".weight"] = 1.0 + _normal(g, (c,), 0.1)
        sd[name + ".bias"] = _normal(g, (c,), 0.1)

    def res(name, ci, co):
        norm(name + ".norm1", ci); conv(name + ".conv1", co, ci, 3)
        norm(name + ".norm2", co); conv(name + ".conv2", co, co, 3)
        if ci != co:
            conv(name + ".nin_shortcut", co, ci, 1)

    def attn(name, c):
        norm(name + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            conv(f"{name}.{n}", c, c, 1)

    n = len(ch_mult)
    in_mult = (1,) + tuple(ch_mult)
    conv("encoder.conv_in", ch, 3, 3)
    bi = ch
    for lvl in range(n):
        bi, bo = ch * in_mult[lvl], ch * ch_mult[lvl]
        for i in range(num_res_blocks):
            res(f"encoder.down.{lvl}.block.{i}", bi, bo)
            bi = bo
        if lvl != n - 1:
            conv(f"encoder.down.{lvl}.downsample.conv", bi, bi, 3)
    res("encoder.mid.block_1", bi, bi); attn("encoder.mid.attn_1", bi); res("encoder.mid.block_2", bi, bi)
    norm("encoder.norm_out", bi); conv("encoder.conv_out", 2 * z_channels, bi, 3)
    bi = ch * ch_mult[-1]
    conv("decoder.conv_in", bi, z_channels, 3)
    res("decoder.mid.block_1", bi, bi); attn("decoder.mid.attn_1", bi); res("decoder.mid.block_2", bi, bi)
    for lvl in reversed(range(n)):
        bo = ch * ch_mult[lvl]
        for i in range(num_res_blocks + 1):
            res(f"decoder.up.{lvl}.block.{i}", bi, bo)
            bi = bo
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", bi, bi, 3)
    norm("decoder.norm_out", bi); conv("decoder.conv_out", 3, bi, 3)
    return {k: v.to(dtype) for k, v in sd.items()}


def vae_inputs(seed: int = 8):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(2, 3, 32, 48, generator=g) * 2 - 1          # ragged-vs-tile shape: W=48 is not a power of two
    noise = torch.randn(2, 16, 16, 24, generator=g)
    return img, noise


class ToyTokenizer:
    """Deterministic offline tokenizer for the inferencer tests (no vocab files offline): an integer word below 1000
    maps to itself, any other whitespace-separated word to a stable id in [0, 900) (NOT Python's salted hash()); ids
    1000-1003 decode to the special-token strings inferencer.gen_text splits on (reference inferencer.py:203-204)."""
    SPECIAL = {1000: "<|im_start|>", 1001: "<|im_end|>", 1002: "<|vision_start|>", 1003: "<|vision_end|>"}

    def encode(self, text):
        ids = []
        for w in text.split():
            if w.isdigit() and int(w) < 1000:
                ids.append(int(w))
            else:
                ids.append(sum(w.encode("utf-8")) * 7 % 900)
        return ids

    def decode(self, ids):
        return " ".join(self.SPECIAL.get(int(i), str(int(i))) for i in ids)


def inferencer_image(seed: int = 3, h: int = 40, w: int = 56):
    """Synthetic RGB input image for the image-edit / understanding flows (smooth gradients + noise), as a PIL image."""
    import numpy as np
    from PIL import Image
    rs = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    img = np.stack([xx, yy, 0.5 + 0.5 * np.sin(6 * xx * yy)], -1) * 200 + rs.randint(0, 55, (h, w, 3))
    return Image.fromarray(img.clip(0, 255).astype(np.uint8))


def train_batch():
    """A hand-built packed TRAINING batch (the format of Bagel.forward, bagel.py:101-150) of two samples:
      sample 0: text (causal) | ViT image (full) | noised VAE image (noise) | text (causal)
      sample 1: clean VAE image (full, timestep -inf => t = 0) | text (causal) | noised VAE image (noise)
    Deterministic; shared by the golden generator and the GPU test (tests/test_gpu_train_forward.py)."""
    g = torch.Generator().manual_seed(31)
    ds, max_lat, max_vit = 16, 8, 8
    from . import bagel_flow as obf
    vit_img = vit_images()[1]                      # 28 x 28 -> 2 x 2 patches
    from . import siglip as osl
    vit_tok = osl.patchify(vit_img, 14)
    lat_shapes = [(4, 4), (2, 4), (2, 2)]                   # latent patch grids (h, w) of the three VAE images
    padded_latent = torch.zeros(3, 16, 8, 8)
    for i, (h, w) in enumerate(lat_shapes):
        padded_latent[i, :, : 2 * h, : 2 * w] = torch.randn(16, 2 * h, 2 * w, generator=g)
    SOI, EOI = 1002, 1003
    text_ids, text_idx, vit_idx, vae_idx, pos, ce_idx, labels = [], [], [], [], [], [], []
    lat_pos, timesteps, mse_idx = [], [], []
    sample_lens, split_lens, attn_modes = [], [], []
    cur = 0

    def add_text(ids, rope, with_loss):
        nonlocal cur
        for k, t in enumerate(ids):
            text_ids.append(t); text_idx.append(cur); pos.append(rope + k)
            if with_loss and k < len(ids) - 1:
                ce_idx.append(cur); labels.append(ids[k + 1])
            cur += 1
        return rope + len(ids)

    def add_vit(rope):
        nonlocal cur
        text_ids.append(SOI); text_idx.append(cur); cur += 1
        vit_idx.extend(range(cur, cur + vit_tok.shape[0])); cur += vit_tok.shape[0]
        text_ids.append(EOI); text_idx.append(cur); cur += 1
        pos.extend([rope] * (vit_tok.shape[0] + 2))
        return rope + 1, vit_tok.shape[0] + 2

    def add_vae(i, rope, t):
        nonlocal cur
        h, w = lat_shapes[i]
        n = h * w
        text_ids.append(SOI); text_idx.append(cur); cur += 1
        vae_idx.extend(range(cur, cur + n))
        if t != float("-inf"):
            mse_idx.extend(range(cur, cur + n))
        cur += n
        text_ids.append(EOI); text_idx.append(cur); cur += 1
        pos.extend([rope] * (n + 2))
        lat_pos.append(obf.flattened_position_ids(h * ds, w * ds, ds, max_lat))
        timesteps.extend([t] * n)
        return rope + 1, n + 2

    # sample 0
    start = cur
    r = add_text([1000, 5, 17, 900, 33, 1001], 0, True); s0 = [6]; m0 = ["causal"]
    r, n = add_vit(r); s0.append(n); m0.append("full")
    r, n = add_vae(0, r, 0.3); s0.append(n); m0.append("noise")
    r = add_text([1000, 8, 100, 4, 1001], r, True); s0.append(5); m0.append("causal")
    sample_lens.append(cur - start); split_lens += s0; attn_modes += m0
    # sample 1
    start = cur
    r, n = add_vae(1, 0, float("-inf")); s1 = [n]; m1 = ["full"]
    r = add_text([1000, 77, 650, 12, 9, 1001], r, False); s1.append(6); m1.append("causal")
    r, n = add_vae(2, r, -0.8); s1.append(n); m1.append("noise")
    sample_lens.append(cur - start); split_lens += s1; attn_modes += m1
    L = cur
    ce_mask = torch.zeros(L, dtype=torch.bool); ce_mask[ce_idx] = True
    mse_mask = torch.zeros(L, dtype=torch.bool); mse_mask[mse_idx] = True
    return dict(
        sequence_length=L, packed_text_ids=torch.tensor(text_ids), packed_text_indexes=torch.tensor(text_idx),
        sample_lens=sample_lens, packed_position_ids=torch.tensor(pos), split_lens=split_lens, attn_modes=attn_modes,
        nested_split_lens=[s0, s1], nested_attn_modes=[m0, m1],
        ce_loss_indexes=ce_mask, packed_label_ids=torch.tensor(labels), packed_vit_tokens=vit_tok,
        packed_vit_token_indexes=torch.tensor(vit_idx),
        packed_vit_position_ids=obf.flattened_position_ids(28, 28, 14, max_vit),
        vit_token_seqlens=torch.tensor([vit_tok.shape[0]], dtype=torch.int), padded_latent=padded_latent,
        patchified_vae_latent_shapes=lat_shapes, packed_latent_position_ids=torch.cat(lat_pos),
        packed_vae_token_indexes=torch.tensor(vae_idx), packed_timesteps=torch.tensor(timesteps),
        mse_loss_indexes=mse_mask)


