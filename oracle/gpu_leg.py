"""TEST INFRASTRUCTURE — the "reference PyTorch path on the GPU" leg (SURVEY.md §8c "On-box GPU oracle").

The oracle (plain torch ops, pinned bit-for-bit to the unmodified reference on CPU by tests/golden/make_golden.py)
runs unchanged on CUDA tensors: nn.Linear -> cuBLAS(Lt) bf16 GEMMs with fp32 accumulation, elementwise ATen kernels,
and — with `fa2()` installed — the reference's real native seam `flash_attn_varlen_func` (qwen2_navit.py:579-588).
That is kernel for kernel what the reference executes on a GPU under autocast. Three legs are used by the drift
test / tool and by bench.py's baseline + parity fields:

  leg "fa2"    oracle on cuda + flash_attn_varlen_func        (the reference as it runs on a GPU)
  leg "sdpa"   oracle on cuda + its own fp32 per-sample SDPA  (the reference as pinned on the CPU, attention shim)
  leg "truth"  oracle under high_precision(): fp32 GEMMs / attention on the same bf16-valued weights

fa2 vs sdpa differ ONLY in the attention kernel's rounding: their distance is the reference's own noise floor.
Nothing in bagel_b200/ imports this module.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import bagel_flow as obf
from . import qwen2_mot as om


def fa2_attention(q, k, v, q_lens, k_lens, causal):
    """flash_attn_varlen_func exactly as the reference calls it (qwen2_navit.py:579-588)."""
    from flash_attn import flash_attn_varlen_func
    dev = q.device
    ql = torch.as_tensor(q_lens, dtype=torch.int32)
    kl = torch.as_tensor(k_lens, dtype=torch.int32)
    cu_q = torch.nn.functional.pad(torch.cumsum(ql, 0, dtype=torch.int32), (1, 0)).to(dev)
    cu_k = torch.nn.functional.pad(torch.cumsum(kl, 0, dtype=torch.int32), (1, 0)).to(dev)
    return flash_attn_varlen_func(q=q.to(torch.bfloat16), k=k.to(torch.bfloat16), v=v.to(torch.bfloat16),
                                  cu_seqlens_q=cu_q, cu_seqlens_k=cu_k, max_seqlen_q=int(ql.max()),
                                  max_seqlen_k=int(kl.max()), causal=bool(causal))


def fa2():
    """with gpu_leg.fa2(): ... -> oracle attention = flash-attn."""
    return om.attention_impl(fa2_attention)


def to_device(d: Dict, dev) -> Dict:
    """Move the tensors of a prepare_* dict to `dev` (the reference does the same before each forward,
    inferencer.py:51-52), leaving lists / ints alone."""
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}


def export_reference_state_dict(model) -> Dict[str, torch.Tensor]:
    """Reference-named state dict (SURVEY.md §8b key schema) rebuilt from a bagel_b200.Bagel's fused kernel layouts —
    so the oracle legs and the product evaluate THE SAME bf16 weights. q/k/v are views of the fused QKV weight;
    gate/up are de-interleaved copies (the product stores them in 128-row gate|up blocks for the SwiGLU epilogue)."""
    lm = model.language_model
    cfg = lm.config
    Hq, Hk, d = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    sd: Dict[str, torch.Tensor] = {}
    P = "language_model."
    sd[P + "model.embed_tokens.weight"] = lm.model.embed_tokens.weight
    for li, layer in enumerate(lm.model.layers):
        p = P + f"model.layers.{li}."
        for sfx, e in (("", layer.und), ("_moe_gen", layer.gen)):
            if e is None:
                continue
            a = p + "self_attn."
            q_w, k_w, v_w = e.wqkv.split([Hq * d, Hk * d, Hk * d], dim=0)
            q_b, k_b, v_b = e.bqkv.split([Hq * d, Hk * d, Hk * d], dim=0)
            sd[a + f"q_proj{sfx}.weight"], sd[a + f"k_proj{sfx}.weight"], sd[a + f"v_proj{sfx}.weight"] = q_w, k_w, v_w
            sd[a + f"q_proj{sfx}.bias"], sd[a + f"k_proj{sfx}.bias"], sd[a + f"v_proj{sfx}.bias"] = q_b, k_b, v_b
            sd[a + f"o_proj{sfx}.weight"] = e.wo
            sd[a + f"q_norm{sfx}.weight"], sd[a + f"k_norm{sfx}.weight"] = e.q_norm, e.k_norm
            I2, K = e.wgu.shape
            gu = e.wgu.view(I2 // 256, 2, 128, K)
            m = p + f"mlp{sfx}."
            sd[m + "gate_proj.weight"] = gu[:, 0].reshape(I2 // 2, K).contiguous()
            sd[m + "up_proj.weight"] = gu[:, 1].reshape(I2 // 2, K).contiguous()
            sd[m + "down_proj.weight"] = e.wd
            sd[p + f"input_layernorm{sfx}.weight"] = e.ln_in
            sd[p + f"post_attention_layernorm{sfx}.weight"] = e.ln_post
    sd[P + "model.norm.weight"] = lm.model.norm
    if lm.model.norm_moe_gen is not None:
        sd[P + "model.norm_moe_gen.weight"] = lm.model.norm_moe_gen
    if lm.lm_head is not None:
        sd[P + "lm_head.weight"] = lm.lm_head.weight
    if getattr(model.config, "visual_gen", False):
        te = model.time_embedder
        sd["time_embedder.mlp.0.weight"], sd["time_embedder.mlp.0.bias"] = te.w0, te.b0
        sd["time_embedder.mlp.2.weight"], sd["time_embedder.mlp.2.bias"] = te.w2, te.b2
        sd["vae2llm.weight"], sd["vae2llm.bias"] = model.vae2llm.weight, model.vae2llm.bias
        sd["llm2vae.weight"], sd["llm2vae.bias"] = model.llm2vae.weight, model.llm2vae.bias
        sd["latent_pos_embed.pos_embed"] = model.latent_pos_embed.pos_embed
    return sd


def flow_config(model) -> obf.FlowConfig:
    cfg = model.language_model.config
    lmc = om.LMConfig(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                      num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                      num_key_value_heads=cfg.num_key_value_heads, vocab_size=cfg.vocab_size,
                      rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta)
    return obf.FlowConfig(lm=lmc, latent_patch_size=model.latent_patch_size, latent_channel=model.latent_channel,
                          vae_downsample=model.config.vae_config.downsample, max_latent_size=model.max_latent_size)


@torch.no_grad()
def reference_contexts(sd, fc: obf.FlowConfig, prompt_ids, new_token_ids, gen_input, cfg_text_input, dev):
    """Prefill exactly as the reference's text->image flow does (prepare_prompts -> forward_cache_update_text for the
    main context; an empty context for the text-CFG branch). Returns (gen_input on dev, main cache, cfg_text branch dict)."""
    L = fc.lm.num_hidden_layers
    B = len(prompt_ids)
    gi_p, kv, rp = obf.prepare_prompts([0] * B, [0] * B, prompt_ids, new_token_ids["bos_token_id"],
                                       new_token_ids["eos_token_id"])
    cache = obf.forward_cache_update_text(sd, fc, om.KVCache(L), **to_device(gi_p, dev))
    gi = to_device(gen_input, dev)
    ct = to_device(cfg_text_input, dev)
    br = dict(packed_position_ids=ct["cfg_packed_position_ids"], packed_query_indexes=ct["cfg_packed_query_indexes"],
              key_values_lens=ct["cfg_key_values_lens"], past_key_values=om.KVCache(L),
              packed_key_value_indexes=ct["cfg_packed_key_value_indexes"])
    return gi, cache, br


@torch.no_grad()
def t2i_reference_run(sd, fc: obf.FlowConfig, prompt_ids, new_token_ids, gen_input, cfg_text_input, dev,
                      x_trace=None, max_steps=None, **sampler_kw):
    """The reference's text->image call order on `dev` with the oracle: prefill (reference_contexts) -> generate_image.
    `gen_input` / `cfg_text_input` are the dicts the product's packers returned (bit-identical to the reference's,
    tests/test_packers.py), so both sides start from the same init noise."""
    gi, cache, br = reference_contexts(sd, fc, prompt_ids, new_token_ids, gen_input, cfg_text_input, dev)
    return obf.generate_image(sd, fc, gi, cache, cfg_text=br, x_trace=x_trace, max_steps=max_steps, **sampler_kw)
