"""-m gpu: dtype mode B — fp32 master weights under autocast, as the reference's eval drivers run the model
(eval/gen/gen_images_mp.py:159-175 + autocast :73; SURVEY.md §8a dtype table): fp32 residual stream, fp32 RMSNorm weights /
outputs, unrounded fp32 RoPE tables, fp32 q/k-norm arithmetic; every nn.Linear bf16 x bf16 -> bf16. Against outputs of
the UNMODIFIED reference run that way (tests/golden/lm_config1.safetensors d64.B.*, mode_b_tiny.safetensors)."""
import os

import pytest
import torch
from safetensors.torch import load_file

import helpers
from oracle import bagel_flow as obf
from oracle import fixtures, qwen2_mot as om
from test_gpu_model import _check

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,cfg,gfile", [("d64", fixtures.TINY_LM, "lm_config1.safetensors"),
                                           ("d128", fixtures.TINY128_LM, "mode_b_tiny.safetensors")])
def test_lm_forward_mode_b(golden_dir, tag, cfg, gfile):
    """und prefill (causal) then a gen forward on the cache; d64 runs the separate q/k-norm+RoPE kernel (flows 2/3),
    d128 the fused QKV-GEMM epilogue with the same flows."""
    from bagel_b200.config import Qwen2Config
    from bagel_b200.qwen2_navit import NaiveCache, Qwen2ForCausalLM
    g = load_file(os.path.join(golden_dir, gfile))
    llm = Qwen2Config(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                      num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                      num_key_value_heads=cfg.num_key_value_heads, rope_theta=cfg.rope_theta, rms_norm_eps=cfg.rms_norm_eps,
                      qk_norm=True, layer_module="Qwen2MoTDecoderLayer")
    lm = Qwen2ForCausalLM(llm, device="cuda", dtype_mode="B")
    sd = fixtures.lm_state_dict(cfg, seed=0, dtype=torch.float32)
    lm.load_state_dict(sd)
    assert lm.model.layers[0].und.ln_in.dtype == torch.float32 and lm.model.layers[0].und.wqkv.dtype == torch.bfloat16
    inp = fixtures.config1_inputs(cfg, dtype=torch.float32)
    cache = NaiveCache(cfg.num_hidden_layers)
    kw_und = dict(query_lens=inp["query_lens"], packed_query_position_ids=inp["und_position_ids"],
                  packed_query_indexes=inp["query_indexes"], key_values_lens=torch.tensor([0], dtype=torch.int32),
                  packed_key_value_indexes=torch.zeros(0, dtype=torch.long), update_past_key_values=True, is_causal=True,
                  mode="und")
    und = lm.forward_inference(packed_query_sequence=inp["x"], past_key_values=cache, **kw_und)
    n = 130
    xg = torch.randn(n, cfg.hidden_size, generator=torch.Generator().manual_seed(5))
    kw_gen = dict(query_lens=torch.tensor([n], dtype=torch.int32), packed_query_position_ids=torch.full((n,), 512, dtype=torch.long),
                  packed_query_indexes=torch.arange(512, 512 + n), key_values_lens=torch.tensor([512], dtype=torch.int32),
                  packed_key_value_indexes=torch.arange(512), update_past_key_values=False, is_causal=False, mode="gen",
                  packed_vae_token_indexes=torch.arange(1, n - 1), packed_text_indexes=torch.tensor([0, n - 1]))
    gen = lm.forward_inference(packed_query_sequence=xg, past_key_values=cache, **kw_gen)
    torch.cuda.synchronize()
    assert und.packed_query_sequence.dtype == torch.float32 and cache.key_cache[0].dtype == torch.bfloat16
    with torch.no_grad(), om.high_precision():
        oc = om.KVCache(cfg.num_hidden_layers)
        t_und, oc = om.lm_forward_inference(sd, cfg, inp["x"], past_key_values=oc, **kw_und)
        t_gen, _ = om.lm_forward_inference(sd, cfg, xg, past_key_values=oc, **kw_gen)
    last = cfg.num_hidden_layers - 1
    pre = f"{tag}.B."
    _check("und hidden", und.packed_query_sequence, g[pre + "und_hidden"], t_und)
    _check("k cache", cache.key_cache[last], g[pre + "k_cache_last"], oc.key_cache[last])
    _check("v cache", cache.value_cache[last], g[pre + "v_cache_last"], oc.value_cache[last])
    _check("gen hidden", gen.packed_query_sequence, g[pre + "gen_hidden"], t_gen)


@pytest.mark.parametrize("tag,cfg", [("d64", fixtures.TINY_LM), ("d128", fixtures.TINY128_LM)])
@pytest.mark.parametrize("name,sT,sI,rt", [("nocfg", 1.0, 1.0, "global"), ("global_img", 4.0, 1.5, "global"),
                                           ("text_channel_img", 4.0, 1.5, "text_channel")])
def test_generate_image_mode_b(golden_dir, tag, cfg, name, sT, sI, rt):
    from bagel_b200.qwen2_navit import NaiveCache
    g = load_file(os.path.join(golden_dir, "mode_b_tiny.safetensors"))
    model = helpers.build_product_bagel(cfg, "cuda", dtype_mode="B")
    assert model.latent_pos_embed.pos_embed.dtype == torch.float32
    tok = helpers.IntTokenizer()

    def ctx(with_text):
        c, kv, rp = NaiveCache(cfg.num_hidden_layers), [0, 0], [0, 0]
        if with_text:
            gi_, kv, rp = model.prepare_prompts(kv, rp, helpers.PROMPTS, tok, helpers.NEW_TOKEN_IDS)
            c = model.forward_cache_update_text(c, **gi_)
        return c, kv, rp

    c_main, kv_m, rp_m = ctx(True)
    c_txt, kv_t, rp_t = ctx(False)
    c_img, kv_i, rp_i = ctx(True)
    _check("prefill k", c_main.key_cache[cfg.num_hidden_layers - 1], g[f"{tag}.prefill.k_cache_last"], None)
    torch.manual_seed(2)
    gi = model.prepare_vae_latent(kv_m, rp_m, helpers.IMAGE_SIZES, helpers.NEW_TOKEN_IDS)
    ct = model.prepare_vae_latent_cfg(kv_t, rp_t, helpers.IMAGE_SIZES)
    ci = model.prepare_vae_latent_cfg(kv_i, rp_i, helpers.IMAGE_SIZES)
    kw = dict(num_timesteps=4, timestep_shift=3.0, cfg_renorm_min=0.0, cfg_renorm_type=rt, cfg_interval=[0.4, 1.0],
              cfg_text_scale=sT, cfg_img_scale=sI)
    lat = model.generate_image(
        past_key_values=c_main, **gi, **kw,
        cfg_text_packed_position_ids=ct["cfg_packed_position_ids"], cfg_text_packed_query_indexes=ct["cfg_packed_query_indexes"],
        cfg_text_key_values_lens=ct["cfg_key_values_lens"], cfg_text_packed_key_value_indexes=ct["cfg_packed_key_value_indexes"],
        cfg_text_past_key_values=c_txt,
        cfg_img_packed_position_ids=ci["cfg_packed_position_ids"], cfg_img_packed_query_indexes=ci["cfg_packed_query_indexes"],
        cfg_img_key_values_lens=ci["cfg_key_values_lens"], cfg_img_packed_key_value_indexes=ci["cfg_packed_key_value_indexes"],
        cfg_img_past_key_values=c_img)
    torch.cuda.synchronize()
    got = torch.cat(lat, 0)
    # exact fp32 evaluation of the same (fp32) weights
    sd = helpers.flow_state_dict(cfg, torch.float32)
    fc = obf.FlowConfig(lm=cfg, max_latent_size=8)

    def octx(with_text):
        c = om.KVCache(cfg.num_hidden_layers)
        if with_text:
            g_, _, _ = obf.prepare_prompts([0, 0], [0, 0], [tok.encode(p) for p in helpers.PROMPTS], 1000, 1001)
            c = obf.forward_cache_update_text(sd, fc, c, **g_)
        return c

    def br(d, cache):
        return dict(packed_position_ids=d["cfg_packed_position_ids"], packed_query_indexes=d["cfg_packed_query_indexes"],
                    key_values_lens=d["cfg_key_values_lens"], past_key_values=cache,
                    packed_key_value_indexes=d["cfg_packed_key_value_indexes"])

    with torch.no_grad(), om.high_precision():
        truth = torch.cat(obf.generate_image(sd, fc, dict(gi), octx(True), cfg_text=br(ct, octx(False)),
                                             cfg_img=br(ci, octx(True)), **kw), 0)
    amp = 1.0 if sT <= 1 else 6.0          # CFG scale 4 x image CFG 1.5 amplifies branch differences (as test_gpu_model.py)
    _check(f"latents[B {tag} {name}]", got, g[f"{tag}.gen.{name}.latents"], truth, max_ulps_of_scale=2.0 * amp)
