"""The oracle (CPU restatement) must reproduce the committed reference outputs bit-for-bit.
Fixtures: tests/golden/*.safetensors, generated from the unmodified reference by tests/golden/make_golden.py."""
import os

import pytest
import torch
from safetensors.torch import load_file

import helpers
from oracle import bagel_flow as obf
from oracle import fixtures, qwen2_mot as om


@pytest.fixture(scope="module")
def g_lm(golden_dir):
    return load_file(os.path.join(golden_dir, "lm_config1.safetensors"))


@pytest.fixture(scope="module")
def g_flow(golden_dir):
    return load_file(os.path.join(golden_dir, "flow_tiny.safetensors"))


@pytest.mark.parametrize("tag,cfg,mode,dtype", [
    ("d64", fixtures.TINY_LM, "A", torch.bfloat16),
    ("d64", fixtures.TINY_LM, "B", torch.float32),
    ("d128", fixtures.TINY128_LM, "A", torch.bfloat16),
])
def test_lm_config1_bit_exact(g_lm, tag, cfg, mode, dtype):
    """BASELINE configs[0]: 2-layer/256-dim MoT forward, seq 512, batch 1, CPU."""
    sd = fixtures.lm_state_dict(cfg, seed=0, dtype=dtype)
    inp = fixtures.config1_inputs(cfg, dtype=dtype)
    with torch.no_grad():
        cache = om.KVCache(cfg.num_hidden_layers)
        h, cache = om.lm_forward_inference(sd, cfg, inp["x"], inp["query_lens"], inp["und_position_ids"],
                                           inp["query_indexes"], cache, torch.tensor([0], dtype=torch.int32),
                                           torch.zeros(0, dtype=torch.long), True, True, "und")
        pre = f"{tag}.{mode}."
        assert torch.equal(h, g_lm[pre + "und_hidden"])
        last = cfg.num_hidden_layers - 1
        assert torch.equal(cache.key_cache[last], g_lm[pre + "k_cache_last"])
        assert torch.equal(cache.value_cache[last], g_lm[pre + "v_cache_last"])
        n = 130
        xg = torch.randn(n, cfg.hidden_size, generator=torch.Generator().manual_seed(5)).to(dtype)
        hg, _ = om.lm_forward_inference(
            sd, cfg, xg, torch.tensor([n], dtype=torch.int32), torch.full((n,), 512, dtype=torch.long),
            torch.arange(512, 512 + n), cache, torch.tensor([512], dtype=torch.int32), torch.arange(512), False,
            False, "gen", torch.arange(1, n - 1), torch.tensor([0, n - 1]))
        assert torch.equal(hg, g_lm[pre + "gen_hidden"])


def _oracle_contexts(sd, fc, cfg):
    tok = helpers.IntTokenizer()
    ids = [tok.encode(p) for p in helpers.PROMPTS]

    def ctx(with_text):
        c = om.KVCache(cfg.num_hidden_layers)
        kv, rp = [0, 0], [0, 0]
        gi = None
        if with_text:
            gi, kv, rp = obf.prepare_prompts(kv, rp, ids, helpers.NEW_TOKEN_IDS["bos_token_id"],
                                             helpers.NEW_TOKEN_IDS["eos_token_id"])
            c = obf.forward_cache_update_text(sd, fc, c, **gi)
        return c, kv, rp, gi

    return ctx


def test_packers_and_prefill_bit_exact(g_flow):
    cfg = fixtures.TINY_LM
    sd = helpers.flow_state_dict(cfg)
    fc = obf.FlowConfig(lm=cfg, max_latent_size=8)
    with torch.no_grad():
        c, kv, rp, gi = _oracle_contexts(sd, fc, cfg)(True)
    for k, v in gi.items():
        assert torch.equal(v, g_flow["prompts." + k]) and v.dtype == g_flow["prompts." + k].dtype, k
    assert kv == g_flow["prefill.kv_lens"].tolist() and rp == g_flow["prefill.ropes"].tolist()
    assert torch.equal(c.key_cache[cfg.num_hidden_layers - 1], g_flow["prefill.k_cache_last"])
    torch.manual_seed(2)
    lat = obf.prepare_vae_latent(fc, kv, rp, helpers.IMAGE_SIZES, helpers.NEW_TOKEN_IDS["start_of_image"],
                                 helpers.NEW_TOKEN_IDS["end_of_image"])
    for k, v in lat.items():
        assert torch.equal(v, g_flow["latent." + k]) and v.dtype == g_flow["latent." + k].dtype, k


@pytest.mark.parametrize("name,sT,sI,rt", [
    ("nocfg", 1.0, 1.0, "global"), ("global", 4.0, 1.0, "global"), ("channel", 4.0, 1.0, "channel"),
    ("global_img", 4.0, 1.5, "global"), ("text_channel_img", 4.0, 1.5, "text_channel"),
    # TaylorSeer step cache (enable_taylorseer=True): 13 evaluations, full at 0-4/7/10, Taylor orders up to 3
    ("taylor_nocfg", 1.0, 1.0, "global"), ("taylor_global_img", 4.0, 1.5, "global"),
    ("taylor_text_channel", 4.0, 1.0, "text_channel")])
def test_generate_image_bit_exact(g_flow, name, sT, sI, rt):
    cfg = fixtures.TINY_LM
    sd = helpers.flow_state_dict(cfg)
    fc = obf.FlowConfig(lm=cfg, max_latent_size=8)
    with torch.no_grad():
        ctx = _oracle_contexts(sd, fc, cfg)
        c_main, kv_m, rp_m, _ = ctx(True)
        c_txt, kv_t, rp_t, _ = ctx(False)
        c_img, kv_i, rp_i, _ = ctx(True)
        torch.manual_seed(2)
        gi = obf.prepare_vae_latent(fc, kv_m, rp_m, helpers.IMAGE_SIZES, helpers.NEW_TOKEN_IDS["start_of_image"],
                                    helpers.NEW_TOKEN_IDS["end_of_image"])

        def br(kv, rp, cache):
            d = obf.prepare_vae_latent_cfg(fc, kv, rp, helpers.IMAGE_SIZES)
            return dict(packed_position_ids=d["cfg_packed_position_ids"],
                        packed_query_indexes=d["cfg_packed_query_indexes"], key_values_lens=d["cfg_key_values_lens"],
                        past_key_values=cache, packed_key_value_indexes=d["cfg_packed_key_value_indexes"])

        taylor = name.startswith("taylor_")
        lat = obf.generate_image(sd, fc, gi, c_main, num_timesteps=14 if taylor else 4, timestep_shift=3.0,
                                 cfg_renorm_min=0.0, cfg_renorm_type=rt, cfg_interval=[0.4, 1.0], cfg_text_scale=sT,
                                 cfg_text=br(kv_t, rp_t, c_txt), cfg_img_scale=sI, cfg_img=br(kv_i, rp_i, c_img),
                                 enable_taylorseer=taylor)
    gold = load_file(os.path.join(os.path.dirname(__file__), "golden", "flow_taylor_tiny.safetensors")) if taylor else g_flow
    assert torch.equal(torch.cat(lat, 0), gold[f"gen.{name}.latents"])


def test_generate_text_greedy_bit_exact(g_flow):
    """Bit-exact greedy token ids (BASELINE north_star) for the oracle vs the reference on CPU."""
    from copy import deepcopy
    cfg = fixtures.TINY_LM
    sd = helpers.flow_state_dict(cfg)
    fc = obf.FlowConfig(lm=cfg, max_latent_size=8)
    with torch.no_grad():
        c, kv, rp, _ = _oracle_contexts(sd, fc, cfg)(True)
        gs = obf.prepare_start_tokens(kv, rp, helpers.NEW_TOKEN_IDS["bos_token_id"])
        for k, v in gs.items():
            assert torch.equal(v, g_flow["start." + k]) and v.dtype == g_flow["start." + k].dtype, k
        trace = []
        toks = obf.generate_text(sd, fc, deepcopy(c), max_length=12, logits_trace=trace, **gs)
    assert torch.equal(toks, g_flow["text.tokens"])
    assert torch.equal(torch.stack(trace, 0), g_flow["text.logits"])


def test_vit_tower_and_context_bit_exact(golden_dir):
    """SigLIP NaViT (head_dim 72) + connector + ViT-context prefill + text prefill on top: oracle == reference."""
    import os
    from oracle import siglip as osl
    g = load_file(os.path.join(golden_dir, "vit_tiny.safetensors"))
    cfg = fixtures.TINY_LM
    tv = fixtures.TINY_VIT
    sd = helpers.vit_flow_state_dict(cfg)
    vc = osl.VitConfig(hidden_size=tv["hidden"], intermediate_size=tv["inter"], num_hidden_layers=tv["layers"],
                       num_attention_heads=tv["heads"])
    gi, kv, rp = osl.prepare_vit_images(vc, 8, [0, 0], [0, 0], fixtures.vit_images(), 1002, 1003)
    for k, v in gi.items():
        assert torch.equal(v, g["vit_in." + k]) and v.dtype == g["vit_in." + k].dtype, k
    assert kv == g["vit_in.kv_lens"].tolist() and rp == g["vit_in.ropes"].tolist()
    with torch.no_grad():
        feats = osl.vit_forward(sd, vc, gi["packed_vit_tokens"], gi["packed_vit_position_ids"], gi["vit_token_seqlens"])
        assert torch.equal(feats, g["vit.features"])
        c = osl.forward_cache_update_vit(sd, cfg, vc, om.KVCache(cfg.num_hidden_layers), **gi)
        gt, _, _ = obf.prepare_prompts(kv, rp, [[5, 17, 900], [8, 8, 100, 4]], 1000, 1001)
        c = obf.forward_cache_update_text(sd, obf.FlowConfig(lm=cfg, max_latent_size=8), c, **gt)
    last = cfg.num_hidden_layers - 1
    assert torch.equal(c.key_cache[last], g["vit.k_cache_last"])
    assert torch.equal(c.value_cache[last], g["vit.v_cache_last"])


def test_vae_bit_exact(golden_dir):
    import os
    from oracle import vae as ov
    g = load_file(os.path.join(golden_dir, "vae_tiny.safetensors"))
    sd = fixtures.vae_state_dict()
    vc = ov.VaeConfig(ch=128, ch_mult=[1, 2], num_res_blocks=1)
    img, noise = fixtures.vae_inputs()
    with torch.no_grad():
        assert torch.equal(ov.encode(sd, vc, img), g["vae.z_mean"])
        assert torch.equal(ov.encode(sd, vc, img, noise), g["vae.z_noise"])
        assert torch.equal(ov.decode(sd, vc, g["vae.z_in"]), g["vae.rec"])


def _vae_ctx_inputs():
    img, _ = fixtures.vae_inputs()
    return [img[0], img[1][:, :24, :32]]


def test_vae_context_prefill_bit_exact(golden_dir):
    """Image-edit context: prepare_vae_images + forward_cache_update_vae (VAE encode with sample=False)."""
    import os
    from oracle import vae as ov
    g = load_file(os.path.join(golden_dir, "vae_tiny.safetensors"))
    cfg = fixtures.TINY_LM
    sd = helpers.flow_state_dict(cfg, max_latent_size=16)
    fc = obf.FlowConfig(lm=cfg, vae_downsample=2, max_latent_size=16)
    vsd = fixtures.vae_state_dict()
    vc = ov.VaeConfig(ch=128, ch_mult=[1, 2], num_res_blocks=1)
    gi, kv, rp = obf.prepare_vae_images(fc, [0, 0], [0, 0], _vae_ctx_inputs(), 1002, 1003)
    for k, v in gi.items():
        if torch.is_tensor(v):
            assert torch.equal(v, g["vae_ctx." + k]), k
    assert kv == g["vae_ctx.kv_lens"].tolist() and rp == g["vae_ctx.ropes"].tolist()
    with torch.no_grad():
        c = obf.forward_cache_update_vae(sd, fc, lambda x: ov.encode(vsd, vc, x), om.KVCache(cfg.num_hidden_layers), **gi)
    last = cfg.num_hidden_layers - 1
    assert torch.equal(c.key_cache[last], g["vae_ctx.k_cache_last"])
    assert torch.equal(c.value_cache[last], g["vae_ctx.v_cache_last"])


# ---------------------------------------------------------------------------------------------------------------------
# round 2 fixtures: decoder-layer variants, SigLIP 2-D RoPE, training forward (all generated by make_golden.py --new-only)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,cfg", [("dense", fixtures.TINY_DENSE_LM), ("moe", fixtures.TINY_MOE_LM)])
def test_dense_and_moe_layers_bit_exact(golden_dir, tag, cfg):
    g = load_file(os.path.join(golden_dir, "lm_variants.safetensors"))
    sd = fixtures.lm_state_dict(cfg, seed=0)
    assert ("model.layers.0.mlp_moe_gen.up_proj.weight" in sd) == (tag == "moe")
    assert "model.layers.0.self_attn.q_proj_moe_gen.weight" not in sd
    inp = fixtures.config1_inputs(cfg)
    with torch.no_grad():
        cache = om.KVCache(cfg.num_hidden_layers)
        h, cache = om.lm_forward_inference(sd, cfg, inp["x"], inp["query_lens"], inp["und_position_ids"], inp["query_indexes"],
                                           cache, torch.tensor([0], dtype=torch.int32), torch.zeros(0, dtype=torch.long),
                                           True, True, "und")
    assert torch.equal(h, g[tag + ".und_hidden"])
    assert torch.equal(cache.key_cache[cfg.num_hidden_layers - 1], g[tag + ".k_cache_last"])


def test_vit_rope_bit_exact(golden_dir):
    from oracle import siglip as osl
    g = load_file(os.path.join(golden_dir, "vit_rope_tiny.safetensors"))
    gin = load_file(os.path.join(golden_dir, "vit_tiny.safetensors"))
    tv = fixtures.TINY_VIT
    vc = osl.VitConfig(hidden_size=tv["hidden"], intermediate_size=tv["inter"], num_hidden_layers=tv["layers"],
                       num_attention_heads=tv["heads"], rope=True, image_size=112)
    with torch.no_grad():
        f = osl.vit_forward(helpers.vit_flow_state_dict(fixtures.TINY_LM), vc, gin["vit_in.packed_vit_tokens"],
                            gin["vit_in.packed_vit_position_ids"], gin["vit_in.vit_token_seqlens"])
    assert torch.equal(f, g["vit_rope.features"])


def test_training_forward_bit_exact(golden_dir):
    """Bagel.forward in train() mode (oracle/train_forward.py) against the reference's losses and hidden states."""
    from oracle import siglip as osl, train_forward as otf
    g = load_file(os.path.join(golden_dir, "train_forward_tiny.safetensors"))
    cfg = fixtures.TINY_LM
    tv = fixtures.TINY_VIT
    b = fixtures.train_batch()
    masks = [otf.prepare_attention_mask_per_sample(s, m) for s, m in zip(b["nested_split_lens"], b["nested_attn_modes"])]
    # mask algebra (data/data_utils.py:72-103): sample 0 = text | ViT (full) | noised VAE (noise) | text
    m0, (a, v, z, t2) = masks[0], b["nested_split_lens"][0]
    assert torch.isinf(m0[0, 1]) and m0[1, 0] == 0                       # causal text
    assert m0[a, a + v - 1] == 0 and m0[a + v - 1, a] == 0               # full image block, both directions
    assert torch.isinf(m0[a + v + z, a + v]).item() and m0[a + v, a + v + z - 1] == 0   # later text cannot see the noise split
    assert m0[a + v + z, a] == 0                                         # ...but sees the ViT image
    vc = osl.VitConfig(hidden_size=tv["hidden"], intermediate_size=tv["inter"], num_hidden_layers=tv["layers"],
                       num_attention_heads=tv["heads"])
    with torch.no_grad():
        o = otf.bagel_forward_train(
            helpers.vit_flow_state_dict(cfg, max_latent_size=8), obf.FlowConfig(lm=cfg, max_latent_size=8), b["sequence_length"],
            b["packed_text_ids"], b["packed_text_indexes"], b["sample_lens"], b["packed_position_ids"], masks, g["train.noise"],
            timestep_shift=1.0, ce_loss_indexes=b["ce_loss_indexes"], packed_label_ids=b["packed_label_ids"],
            vit=(vc, b["packed_vit_tokens"], b["packed_vit_token_indexes"], b["packed_vit_position_ids"], b["vit_token_seqlens"]),
            padded_latent=b["padded_latent"], patchified_vae_latent_shapes=b["patchified_vae_latent_shapes"],
            packed_latent_position_ids=b["packed_latent_position_ids"], packed_vae_token_indexes=b["packed_vae_token_indexes"],
            packed_timesteps=b["packed_timesteps"], mse_loss_indexes=b["mse_loss_indexes"])
    assert torch.equal(o["last_hidden_state"], g["train.last_hidden_state"])
    assert torch.equal(o["mse"], g["train.mse"]) and torch.equal(o["ce"], g["train.ce"])


def test_mode_b_flow_bit_exact(golden_dir):
    """fp32 master weights under autocast (eval drivers): text prefill + 3-evaluation generate_image, d64."""
    g = load_file(os.path.join(golden_dir, "mode_b_tiny.safetensors"))
    cfg = fixtures.TINY_LM
    sd = helpers.flow_state_dict(cfg, torch.float32)
    fc = obf.FlowConfig(lm=cfg, max_latent_size=8)
    tok = helpers.IntTokenizer()
    with torch.no_grad():
        gp, kv, rp = obf.prepare_prompts([0, 0], [0, 0], [tok.encode(p) for p in helpers.PROMPTS], 1000, 1001)
        cache = obf.forward_cache_update_text(sd, fc, om.KVCache(cfg.num_hidden_layers), **gp)
        assert torch.equal(cache.key_cache[cfg.num_hidden_layers - 1], g["d64.prefill.k_cache_last"])
        torch.manual_seed(2)
        gi = obf.prepare_vae_latent(fc, kv, rp, helpers.IMAGE_SIZES, 1002, 1003)
        lat = obf.generate_image(sd, fc, gi, cache, num_timesteps=4, timestep_shift=3.0, cfg_renorm_type="global",
                                 cfg_interval=[0.4, 1.0])
    assert torch.equal(torch.cat(lat, 0), g["d64.gen.nocfg.latents"])
