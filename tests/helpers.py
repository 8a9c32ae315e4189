"""Shared builders for the tests: tiny synthetic BAGEL (weights from oracle/fixtures.py), int tokenizer."""
from __future__ import annotations

import torch

from oracle import bagel_flow as obf
from oracle import fixtures

NEW_TOKEN_IDS = dict(bos_token_id=1000, eos_token_id=1001, start_of_image=1002, end_of_image=1003)
PROMPTS = ["5 17 900 33 2", "8 8 100 4 77 650 12"]
IMAGE_SIZES = [(64, 64), (64, 96)]


class IntTokenizer:
    def encode(self, prompt):
        return [int(t) for t in prompt.split()]


def flow_state_dict(cfg=fixtures.TINY_LM, dtype=torch.bfloat16, max_latent_size=8):
    sd = {"language_model." + k: v for k, v in fixtures.lm_state_dict(cfg, seed=0, dtype=dtype).items()}
    sd.update(fixtures.bagel_extra_state_dict(cfg.hidden_size, seed=1, dtype=dtype))
    sd["latent_pos_embed.pos_embed"] = obf.sincos_2d_table(cfg.hidden_size, max_latent_size).to(dtype)
    return sd


def tiny_vae(device="cuda"):
    from bagel_b200.autoencoder import AutoEncoder
    from bagel_b200.config import AutoEncoderParams
    ae = AutoEncoder(AutoEncoderParams(resolution=32, downsample=2, ch=128, ch_mult=[1, 2], num_res_blocks=1,
                                       z_channels=16), device)
    ae.load_state_dict(fixtures.vae_state_dict())
    return ae


def build_product_bagel(cfg=fixtures.TINY_LM, device="cuda", max_latent_size=8, load=True, vae_downsample=8,
                        dtype_mode="A"):
    """bagel_b200.Bagel for the tiny config (device='cpu' only exercises host logic: packers, config).
    dtype_mode "B": fp32 master weights (the state dict is then generated in fp32)."""
    from bagel_b200.bagel import Bagel
    from bagel_b200.config import AutoEncoderParams, BagelConfig, Qwen2Config
    from bagel_b200.qwen2_navit import Qwen2ForCausalLM

    llm = Qwen2Config(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                      num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                      num_key_value_heads=cfg.num_key_value_heads, rope_theta=cfg.rope_theta,
                      rms_norm_eps=cfg.rms_norm_eps, qk_norm=True, layer_module="Qwen2MoTDecoderLayer")
    bcfg = BagelConfig(visual_gen=True, visual_und=False, llm_config=llm, vit_config=None,
                       vae_config=AutoEncoderParams(downsample=vae_downsample), latent_patch_size=2,
                       max_latent_size=max_latent_size)
    lm = Qwen2ForCausalLM(llm, device=device, dtype_mode=dtype_mode)
    model = Bagel(lm, None, bcfg)
    if load:
        model.load_state_dict(flow_state_dict(cfg, torch.float32 if dtype_mode == "B" else torch.bfloat16,
                                              max_latent_size=max_latent_size))
    return model


def vit_flow_state_dict(cfg=fixtures.TINY_LM, dtype=torch.bfloat16, max_latent_size=8):
    tv = fixtures.TINY_VIT
    sd = flow_state_dict(cfg, dtype, max_latent_size)
    sd.update(fixtures.vit_state_dict(tv["hidden"], tv["inter"], tv["layers"], tv["heads"], cfg.hidden_size, dtype=dtype))
    sd["vit_pos_embed.pos_embed"] = obf.sincos_2d_table(cfg.hidden_size, 8).to(dtype)
    return sd


def build_product_bagel_with_vit(cfg=fixtures.TINY_LM, device="cuda", load=True, max_latent_size=8, vae_downsample=8):
    from bagel_b200.bagel import Bagel
    from bagel_b200.config import AutoEncoderParams, BagelConfig, Qwen2Config, SiglipVisionConfig
    from bagel_b200.qwen2_navit import Qwen2ForCausalLM
    from bagel_b200.siglip_navit import SiglipVisionModel

    tv = fixtures.TINY_VIT
    llm = Qwen2Config(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                      num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                      num_key_value_heads=cfg.num_key_value_heads, rope_theta=cfg.rope_theta,
                      rms_norm_eps=cfg.rms_norm_eps, qk_norm=True, layer_module="Qwen2MoTDecoderLayer")
    vcfg = SiglipVisionConfig(hidden_size=tv["hidden"], intermediate_size=tv["inter"], num_hidden_layers=tv["layers"],
                              num_attention_heads=tv["heads"], num_channels=3, image_size=112, patch_size=14, rope=False)
    bcfg = BagelConfig(visual_gen=True, visual_und=True, llm_config=llm, vit_config=vcfg,
                       vae_config=AutoEncoderParams(downsample=vae_downsample), latent_patch_size=2,
                       max_latent_size=max_latent_size, vit_max_num_patch_per_side=8)
    model = Bagel(Qwen2ForCausalLM(llm, device=device), SiglipVisionModel(vcfg, device=device), bcfg)
    if load:
        model.load_state_dict(vit_flow_state_dict(cfg, max_latent_size=max_latent_size))
    return model
