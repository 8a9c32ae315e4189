"""Synthetic (random-init) BAGEL-7B-MoT for benchmarks and smoke tests: there is no network for checkpoints, so
weights are drawn on the device with the reference's init recipe (normal, std = initializer_range;
modeling/qwen2/modeling_qwen2.py:563-572), llm2vae given non-zero weights (the reference zero-inits it,
bagel.py:96-99, which would make the velocity constant — SURVEY.md A.10), prompts are random token ids.
Shapes are the public BAGEL-7B-MoT llm_config.json (Qwen2.5-7B): SURVEY.md §8.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from .bagel import Bagel
from .config import AutoEncoderParams, BagelConfig, Qwen2Config
from .qwen2_navit import Qwen2ForCausalLM, _Embedding, _ExpertWeights, _Linear

BF16 = torch.bfloat16

BAGEL_7B_LLM = dict(vocab_size=152064, hidden_size=3584, intermediate_size=18944, num_hidden_layers=28,
                    num_attention_heads=28, num_key_value_heads=4, max_position_embeddings=32768,
                    rms_norm_eps=1e-6, rope_theta=1000000.0, qk_norm=True, tie_word_embeddings=False,
                    layer_module="Qwen2MoTDecoderLayer")

NEW_TOKEN_IDS = dict(bos_token_id=151644, eos_token_id=151645, start_of_image=151652, end_of_image=151653)


class RandomIdTokenizer:
    """encode("<n>") -> n deterministic pseudo-random ids in [0, 151643) (no vocab files offline)."""

    def __init__(self, seed: int = 1, vocab: int = 151643):
        self.gen = torch.Generator().manual_seed(seed)
        self.vocab = vocab

    def encode(self, prompt: str) -> List[int]:
        n = int(prompt)
        return torch.randint(0, self.vocab, (n,), generator=self.gen).tolist()


def _randn(shape, std, gen, device):
    return (torch.randn(shape, generator=gen, device=device, dtype=torch.float32) * std).to(BF16)


def build_random_bagel(llm_kwargs: Dict = None, device="cuda", seed: int = 0, max_latent_size: int = 64,
                       num_layers: int = None) -> Bagel:
    """Random-init MoT model directly in the kernels' fused layouts (no intermediate fp32 state dict: the 7B
    model is 28 GB in bf16)."""
    kw = dict(BAGEL_7B_LLM if llm_kwargs is None else llm_kwargs)
    if num_layers is not None:
        kw["num_hidden_layers"] = num_layers
    llm = Qwen2Config(**kw)
    bcfg = BagelConfig(visual_gen=True, visual_und=False, llm_config=llm, vit_config=None,
                       vae_config=AutoEncoderParams(), latent_patch_size=2, max_latent_size=max_latent_size)
    lm = Qwen2ForCausalLM(llm, device=device)
    dev = lm.device
    g = torch.Generator(device=dev).manual_seed(seed)
    std = llm.initializer_range
    H, I, d = llm.hidden_size, llm.intermediate_size, llm.head_dim
    Hq, Hk = llm.num_attention_heads, llm.num_key_value_heads
    lm.model.embed_tokens = _Embedding(_randn((llm.vocab_size, H), std, g, dev))
    for layer in lm.model.layers:
        for tgt in ("und", "gen"):
            e = _ExpertWeights()
            e.wqkv = _randn(((Hq + 2 * Hk) * d, H), std, g, dev)
            e.bqkv = _randn(((Hq + 2 * Hk) * d,), std, g, dev)
            e.wo = _randn((H, Hq * d), std, g, dev)
            e.wgu = _randn((2 * I, H), std, g, dev)       # already in the interleaved gate|up layout
            e.wd = _randn((H, I), std, g, dev)
            e.ln_in = (1.0 + _randn((H,), 0.02, g, dev).float()).to(BF16)
            e.ln_post = (1.0 + _randn((H,), 0.02, g, dev).float()).to(BF16)
            e.q_norm = (1.0 + _randn((d,), 0.02, g, dev).float()).to(BF16)
            e.k_norm = (1.0 + _randn((d,), 0.02, g, dev).float()).to(BF16)
            setattr(layer, tgt, e)
    lm.model.norm = (1.0 + _randn((H,), 0.02, g, dev).float()).to(BF16)
    lm.model.norm_moe_gen = (1.0 + _randn((H,), 0.02, g, dev).float()).to(BF16)
    lm.lm_head = _Linear(_randn((llm.vocab_size, H), std, g, dev))
    model = Bagel(lm, None, bcfg)
    sd = {
        "time_embedder.mlp.0.weight": _randn((H, 256), std, g, dev), "time_embedder.mlp.0.bias": _randn((H,), std, g, dev),
        "time_embedder.mlp.2.weight": _randn((H, H), std, g, dev), "time_embedder.mlp.2.bias": _randn((H,), std, g, dev),
        "vae2llm.weight": _randn((H, model.patch_latent_dim), std, g, dev), "vae2llm.bias": _randn((H,), std, g, dev),
        "llm2vae.weight": _randn((model.patch_latent_dim, H), std, g, dev),
        "llm2vae.bias": _randn((model.patch_latent_dim,), std, g, dev),
    }
    model.time_embedder.load(sd, "time_embedder.", dev)
    model.vae2llm.load(sd, "vae2llm.", dev)
    model.llm2vae.load(sd, "llm2vae.", dev)
    return model


def t2i_inputs(model: Bagel, batch: int, image_size=(1024, 1024), prompt_tokens: int = 64, seed: int = 1,
               noise_seed: int = 2):
    """BASELINE configs[1] inputs (SURVEY.md §8d cfg 2): per sample 64 random ids + bos/eos as context, prefilled
    through forward_cache_update_text; returns (gen_input, cfg_text_input, contexts)."""
    from .qwen2_navit import NaiveCache
    L = model.config.llm_config.num_hidden_layers
    tok = RandomIdTokenizer(seed)
    kv0, rp0 = [0] * batch, [0] * batch
    gi, kv, rp = model.prepare_prompts(kv0, rp0, [str(prompt_tokens)] * batch, tok, NEW_TOKEN_IDS)
    cache = model.forward_cache_update_text(NaiveCache(L), **gi)
    torch.manual_seed(noise_seed)
    gen_input = model.prepare_vae_latent(kv, rp, [image_size] * batch, NEW_TOKEN_IDS)
    cfg_text = model.prepare_vae_latent_cfg(kv0, rp0, [image_size] * batch)
    return gen_input, cfg_text, dict(main=cache, cfg_text=NaiveCache(L), kv_lens=kv, ropes=rp)
