"""Generate the committed golden fixtures by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no reference tree):
    python tests/golden/make_golden.py
Writes tests/golden/*.safetensors. The reference ships no tests / golden vectors, so these are
the pins: every tensor below is produced by the reference's own classes (with the harness shims of
oracle/ref_shims.py), from the deterministic synthetic weights/inputs of oracle/fixtures.py. While generating,
the script also asserts that the oracle restatement reproduces the reference bit-for-bit.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import torch
from safetensors.torch import save_file

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import bagel_flow as obf  # noqa: E402
from oracle import fixtures, qwen2_mot as om, ref_shims  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
NEW_TOKEN_IDS = dict(bos_token_id=1000, eos_token_id=1001, start_of_image=1002, end_of_image=1003)


class IntTokenizer:
    """Prompts are strings of space-separated token ids (no vocab files offline)."""

    def encode(self, prompt):
        return [int(t) for t in prompt.split()]


def ref_lm(ns, cfg, sd, dtype):
    rcfg = ref_shims.make_llm_config(
        ns, vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
        num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
        num_key_value_heads=cfg.num_key_value_heads, rope_theta=cfg.rope_theta, max_position_embeddings=32768,
        layer_module=cfg.layer_module)
    lm = ref_shims.cast_parameters(ns.qwen2_navit.Qwen2ForCausalLM(rcfg).eval(), dtype)
    lm.load_state_dict(sd, strict=True)
    return lm, rcfg


def golden_lm_config1(ns):
    """BASELINE configs[0]: 2-layer / 256-dim MoT forward, seq 512, batch 1 (und prefill, then gen on top)."""
    out = {}
    for tag, cfg in (("d64", fixtures.TINY_LM), ("d128", fixtures.TINY128_LM)):
        for mode, dtype in (("A", torch.bfloat16), ("B", torch.float32)):
            if tag == "d128" and mode == "B":
                continue
            sd = fixtures.lm_state_dict(cfg, seed=0, dtype=dtype)
            lm, _ = ref_lm(ns, cfg, sd, dtype)
            inp = fixtures.config1_inputs(cfg, dtype=dtype)
            n = 130
            xg = torch.randn(n, cfg.hidden_size, generator=torch.Generator().manual_seed(5)).to(dtype)
            kw = dict(query_lens=torch.tensor([n], dtype=torch.int32),
                      packed_query_position_ids=torch.full((n,), 512, dtype=torch.long),
                      packed_query_indexes=torch.arange(512, 512 + n),
                      key_values_lens=torch.tensor([512], dtype=torch.int32),
                      packed_key_value_indexes=torch.arange(512), update_past_key_values=False, is_causal=False,
                      mode="gen", packed_vae_token_indexes=torch.arange(1, n - 1),
                      packed_text_indexes=torch.tensor([0, n - 1]))
            with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
                cache = ns.qwen2_navit.NaiveCache(cfg.num_hidden_layers)
                und = lm.forward_inference(
                    packed_query_sequence=inp["x"], query_lens=inp["query_lens"],
                    packed_query_position_ids=inp["und_position_ids"], packed_query_indexes=inp["query_indexes"],
                    past_key_values=cache, key_values_lens=torch.tensor([0], dtype=torch.int32),
                    packed_key_value_indexes=torch.zeros(0, dtype=torch.long), update_past_key_values=True,
                    is_causal=True, mode="und")
                gen = lm.forward_inference(packed_query_sequence=xg, past_key_values=cache, **kw)
            with torch.no_grad():
                oc = om.KVCache(cfg.num_hidden_layers)
                oh, oc = om.lm_forward_inference(sd, cfg, inp["x"], inp["query_lens"], inp["und_position_ids"],
                                                 inp["query_indexes"], oc, torch.tensor([0], dtype=torch.int32),
                                                 torch.zeros(0, dtype=torch.long), True, True, "und")
                og, _ = om.lm_forward_inference(sd, cfg, xg, past_key_values=oc, **kw)
            assert torch.equal(oh, und.packed_query_sequence) and torch.equal(og, gen.packed_query_sequence), \
                f"oracle != reference ({tag}, mode {mode})"
            for li in range(cfg.num_hidden_layers):
                assert torch.equal(oc.key_cache[li], cache.key_cache[li])
            pre = f"{tag}.{mode}."
            out[pre + "und_hidden"] = und.packed_query_sequence.contiguous()
            out[pre + "gen_hidden"] = gen.packed_query_sequence.contiguous()
            out[pre + "k_cache_last"] = cache.key_cache[cfg.num_hidden_layers - 1].contiguous()
            out[pre + "v_cache_last"] = cache.value_cache[cfg.num_hidden_layers - 1].contiguous()
            print(f"lm_config1 {tag} mode {mode}: oracle == reference (bit-exact)")
    save_file(out, os.path.join(OUT, "lm_config1.safetensors"))


def build_ref_bagel(ns, cfg, sd_all, dtype, max_latent_size=8):
    lm_sd = {k[len("language_model."):]: v for k, v in sd_all.items() if k.startswith("language_model.")}
    lm, rcfg = ref_lm(ns, cfg, lm_sd, dtype)
    vae_cfg = SimpleNamespace(downsample=8, z_channels=16)
    bcfg = ns.bagel.BagelConfig(visual_gen=True, visual_und=False, llm_config=rcfg, vit_config=None,
                                vae_config=vae_cfg, latent_patch_size=2, max_latent_size=max_latent_size)
    model = ns.bagel.Bagel(lm, None, bcfg).eval()
    ref_shims.cast_parameters(model, dtype)
    missing = model.load_state_dict(sd_all, strict=False)
    assert not missing.unexpected_keys, missing
    assert all("pos_embed" in k for k in missing.missing_keys), missing
    return model


def flow_state_dict(cfg, dtype, max_latent_size=8):
    sd = {"language_model." + k: v for k, v in fixtures.lm_state_dict(cfg, seed=0, dtype=dtype).items()}
    sd.update(fixtures.bagel_extra_state_dict(cfg.hidden_size, seed=1, dtype=dtype))
    return sd


def golden_flow(ns):
    """Packers + text prefill + generate_image (4 timesteps = 3 evals, B=2 ragged images, CFG variants)."""
    cfg = fixtures.TINY_LM
    dtype = torch.bfloat16
    sd = flow_state_dict(cfg, dtype)
    model = build_ref_bagel(ns, cfg, sd, dtype)
    sd_full = dict(sd)
    sd_full["latent_pos_embed.pos_embed"] = model.latent_pos_embed.pos_embed.data.clone()
    fc = obf.FlowConfig(lm=cfg, max_latent_size=8)
    assert torch.equal(obf.sincos_2d_table(cfg.hidden_size, 8).to(dtype), sd_full["latent_pos_embed.pos_embed"])
    tok = IntTokenizer()
    prompts = ["5 17 900 33 2", "8 8 100 4 77 650 12"]
    prompt_ids = [tok.encode(p) for p in prompts]
    image_sizes = [(64, 64), (64, 96)]
    out = {}

    def ctx(with_text):
        """(cache, kv_lens, ropes) for the reference and the oracle."""
        rc = ns.qwen2_navit.NaiveCache(cfg.num_hidden_layers)
        oc = om.KVCache(cfg.num_hidden_layers)
        kv, rp = [0, 0], [0, 0]
        if with_text:
            gi, kv2, rp2 = model.prepare_prompts(kv, rp, prompts, tok, NEW_TOKEN_IDS)
            ogi, okv, orp = obf.prepare_prompts(kv, rp, prompt_ids, NEW_TOKEN_IDS["bos_token_id"],
                                                NEW_TOKEN_IDS["eos_token_id"])
            assert kv2 == okv and rp2 == orp
            for k in gi:
                assert torch.equal(gi[k], ogi[k]) and gi[k].dtype == ogi[k].dtype, k
                out["prompts." + k] = gi[k].clone()
            with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
                rc = model.forward_cache_update_text(rc, **gi)
            with torch.no_grad():
                oc = obf.forward_cache_update_text(sd_full, fc, oc, **ogi)
            for li in range(cfg.num_hidden_layers):
                assert torch.equal(rc.key_cache[li], oc.key_cache[li])
            kv, rp = kv2, rp2
        return rc, oc, kv, rp

    rc_main, oc_main, kv_main, rp_main = ctx(True)
    rc_txt, oc_txt, kv_txt, rp_txt = ctx(False)      # text-dropped branch: empty context
    rc_img, oc_img, kv_img, rp_img = ctx(True)       # "image-dropped" branch of a T2I call = text only
    out["prefill.k_cache_last"] = rc_main.key_cache[cfg.num_hidden_layers - 1].contiguous()
    out["prefill.kv_lens"] = torch.tensor(kv_main)
    out["prefill.ropes"] = torch.tensor(rp_main)

    torch.manual_seed(2)
    gi = model.prepare_vae_latent(kv_main, rp_main, image_sizes, NEW_TOKEN_IDS)
    torch.manual_seed(2)
    ogi = obf.prepare_vae_latent(fc, kv_main, rp_main, image_sizes, NEW_TOKEN_IDS["start_of_image"],
                                 NEW_TOKEN_IDS["end_of_image"])
    for k in gi:
        assert torch.equal(gi[k], ogi[k]) and gi[k].dtype == ogi[k].dtype, k
        out["latent." + k] = gi[k].clone()
    cfg_t = model.prepare_vae_latent_cfg(kv_txt, rp_txt, image_sizes)
    cfg_i = model.prepare_vae_latent_cfg(kv_img, rp_img, image_sizes)
    ocfg_t = obf.prepare_vae_latent_cfg(fc, kv_txt, rp_txt, image_sizes)
    for k in cfg_t:
        assert torch.equal(cfg_t[k], ocfg_t[k]), k
        out["cfg_text." + k] = cfg_t[k].clone()
        out["cfg_img." + k] = cfg_i[k].clone()

    def obranch(d, cache):
        return dict(packed_position_ids=d["cfg_packed_position_ids"], packed_query_indexes=d["cfg_packed_query_indexes"],
                    key_values_lens=d["cfg_key_values_lens"], past_key_values=cache,
                    packed_key_value_indexes=d["cfg_packed_key_value_indexes"])

    variants = [("nocfg", 1.0, 1.0, "global"), ("global", 4.0, 1.0, "global"), ("channel", 4.0, 1.0, "channel"),
                ("global_img", 4.0, 1.5, "global"), ("text_channel_img", 4.0, 1.5, "text_channel")]
    for name, sT, sI, rt in variants:
        kwargs = dict(num_timesteps=4, timestep_shift=3.0, cfg_renorm_min=0.0, cfg_renorm_type=rt,
                      cfg_interval=[0.4, 1.0], cfg_text_scale=sT, cfg_img_scale=sI)
        ref_kw = dict(kwargs)
        ref_kw.update(
            cfg_text_packed_position_ids=cfg_t["cfg_packed_position_ids"],
            cfg_text_packed_query_indexes=cfg_t["cfg_packed_query_indexes"],
            cfg_text_key_values_lens=cfg_t["cfg_key_values_lens"],
            cfg_text_packed_key_value_indexes=cfg_t["cfg_packed_key_value_indexes"],
            cfg_text_past_key_values=rc_txt,
            cfg_img_packed_position_ids=cfg_i["cfg_packed_position_ids"],
            cfg_img_packed_query_indexes=cfg_i["cfg_packed_query_indexes"],
            cfg_img_key_values_lens=cfg_i["cfg_key_values_lens"],
            cfg_img_packed_key_value_indexes=cfg_i["cfg_packed_key_value_indexes"],
            cfg_img_past_key_values=rc_img)
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            lat = model.generate_image(past_key_values=rc_main, **gi, **ref_kw)
        with torch.no_grad():
            olat = obf.generate_image(sd_full, fc, ogi, oc_main, cfg_text=obranch(cfg_t, oc_txt),
                                      cfg_img=obranch(cfg_i, oc_img), **kwargs)
        for a, b in zip(lat, olat):
            assert torch.equal(a, b), f"oracle generate_image != reference ({name})"
        out[f"gen.{name}.latents"] = torch.cat(lat, dim=0).contiguous()
        print(f"generate_image[{name}]: oracle == reference (bit-exact); |x| mean {float(torch.cat(lat).abs().mean()):.4f}")
    # ---- TaylorSeer step cache (enable_taylorseer=True, bagel.py:678-684; cache_utils/taylorseer.py): 14 timesteps =
    # 13 evaluations -> full steps 0-4, 7, 10 (Taylor orders 1, 2, 3), extrapolated steps 5, 6, 8, 9, 11, 12; the CFG
    # branches stop at the cfg_interval boundary with their own step counters ----
    out_ts = {}
    for name, sT, sI, rt in [("taylor_nocfg", 1.0, 1.0, "global"), ("taylor_global_img", 4.0, 1.5, "global"),
                             ("taylor_text_channel", 4.0, 1.0, "text_channel")]:
        kwargs = dict(num_timesteps=14, timestep_shift=3.0, cfg_renorm_min=0.0, cfg_renorm_type=rt,
                      cfg_interval=[0.4, 1.0], cfg_text_scale=sT, cfg_img_scale=sI, enable_taylorseer=True)
        ref_kw = dict(kwargs)
        ref_kw.update(
            cfg_text_packed_position_ids=cfg_t["cfg_packed_position_ids"],
            cfg_text_packed_query_indexes=cfg_t["cfg_packed_query_indexes"],
            cfg_text_key_values_lens=cfg_t["cfg_key_values_lens"],
            cfg_text_packed_key_value_indexes=cfg_t["cfg_packed_key_value_indexes"],
            cfg_text_past_key_values=rc_txt,
            cfg_img_packed_position_ids=cfg_i["cfg_packed_position_ids"],
            cfg_img_packed_query_indexes=cfg_i["cfg_packed_query_indexes"],
            cfg_img_key_values_lens=cfg_i["cfg_key_values_lens"],
            cfg_img_packed_key_value_indexes=cfg_i["cfg_packed_key_value_indexes"],
            cfg_img_past_key_values=rc_img)
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            lat = model.generate_image(past_key_values=rc_main, **gi, **ref_kw)
        with torch.no_grad():
            olat = obf.generate_image(sd_full, fc, ogi, oc_main, cfg_text=obranch(cfg_t, oc_txt),
                                      cfg_img=obranch(cfg_i, oc_img), **kwargs)
        for a, b in zip(lat, olat):
            assert torch.equal(a, b), f"oracle generate_image != reference ({name})"
        out_ts[f"gen.{name}.latents"] = torch.cat(lat, dim=0).contiguous()
        print(f"generate_image[{name}]: oracle == reference (bit-exact); |x| mean {float(torch.cat(lat).abs().mean()):.4f}")
    save_file(out_ts, os.path.join(OUT, "flow_taylor_tiny.safetensors"))
    # the reference leaves the flag set on the model AND on every decoder layer after such a call (bagel.py:681,
    # qwen2_navit.py:1060); a later non-TaylorSeer forward in the same process would then read the stale cache
    model.language_model.model.enable_taylorseer = False
    for layer in model.language_model.model.layers:
        layer.enable_taylorseer = False
    # ---- greedy text decode on top of the text context (bagel.py:930-1000), 12 steps, B=2 ----
    from copy import deepcopy
    gs = model.prepare_start_tokens(kv_main, rp_main, NEW_TOKEN_IDS)
    ogs = obf.prepare_start_tokens(kv_main, rp_main, NEW_TOKEN_IDS["bos_token_id"])
    for k in gs:
        assert torch.equal(gs[k], ogs[k]) and gs[k].dtype == ogs[k].dtype, k
        out["start." + k] = gs[k].clone()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        toks = model.generate_text(past_key_values=deepcopy(rc_main), max_length=12, do_sample=False, **gs)
    trace = []
    with torch.no_grad():
        otoks = obf.generate_text(sd_full, fc, deepcopy(oc_main), max_length=12, logits_trace=trace, **ogs)
    assert torch.equal(toks, otoks), "oracle generate_text != reference"
    out["text.tokens"] = toks.contiguous()
    out["text.logits"] = torch.stack(trace, 0).contiguous()      # [steps, B, vocab] bf16 (oracle == reference path)
    top2 = torch.stack(trace, 0).float().topk(2, dim=-1).values
    print("generate_text: oracle == reference (bit-exact); tokens", toks[:, 0].tolist(),
          "min top-1/top-2 logit margin", float((top2[..., 0] - top2[..., 1]).min()))
    save_file(out, os.path.join(OUT, "flow_tiny.safetensors"))


def golden_vit(ns):
    """SigLIP NaViT tower (head_dim 72) + connector + ViT context prefill, then a text prefill on top."""
    from oracle import siglip as osl
    cfg = fixtures.TINY_LM
    dtype = torch.bfloat16
    tv = fixtures.TINY_VIT
    sd = flow_state_dict(cfg, dtype)
    sd.update(fixtures.vit_state_dict(tv["hidden"], tv["inter"], tv["layers"], tv["heads"], cfg.hidden_size, dtype=dtype))
    lm_sd = {k[len("language_model."):]: v for k, v in sd.items() if k.startswith("language_model.")}
    lm, rcfg = ref_lm(ns, cfg, lm_sd, dtype)
    sn = ns.siglip_navit
    vcfg = sn.SiglipVisionConfig(hidden_size=tv["hidden"], intermediate_size=tv["inter"], num_hidden_layers=tv["layers"],
                                 num_attention_heads=tv["heads"], num_channels=3, image_size=112, patch_size=14, rope=False)
    vit = sn.SiglipVisionModel(vcfg)
    vit.vision_model.embeddings.convert_conv2d_to_linear(vcfg)
    bcfg = ns.bagel.BagelConfig(visual_gen=True, visual_und=True, llm_config=rcfg, vit_config=vcfg,
                                vae_config=SimpleNamespace(downsample=8, z_channels=16), latent_patch_size=2,
                                max_latent_size=8, vit_max_num_patch_per_side=8)
    model = ns.bagel.Bagel(lm, vit, bcfg).eval()
    ref_shims.cast_parameters(model, dtype)
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all("pos_embed" in k for k in missing.missing_keys), missing
    sd_full = dict(sd)
    sd_full["latent_pos_embed.pos_embed"] = model.latent_pos_embed.pos_embed.data.clone()
    sd_full["vit_pos_embed.pos_embed"] = model.vit_pos_embed.pos_embed.data.clone()
    assert torch.equal(obf.sincos_2d_table(cfg.hidden_size, 8).to(dtype), sd_full["vit_pos_embed.pos_embed"])
    vc = osl.VitConfig(hidden_size=tv["hidden"], intermediate_size=tv["inter"], num_hidden_layers=tv["layers"],
                       num_attention_heads=tv["heads"])
    images = fixtures.vit_images()
    out = {}
    gi, kv, rp = model.prepare_vit_images([0, 0], [0, 0], images, lambda im: im, NEW_TOKEN_IDS)
    ogi, okv, orp = osl.prepare_vit_images(vc, 8, [0, 0], [0, 0], images, NEW_TOKEN_IDS["start_of_image"],
                                           NEW_TOKEN_IDS["end_of_image"])
    assert kv == okv and rp == orp
    for k in gi:
        assert torch.equal(gi[k], ogi[k]) and gi[k].dtype == ogi[k].dtype, k
        out["vit_in." + k] = gi[k].clone()
    out["vit_in.kv_lens"] = torch.tensor(kv)
    out["vit_in.ropes"] = torch.tensor(rp)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        cu = torch.nn.functional.pad(torch.cumsum(gi["vit_token_seqlens"], 0), (1, 0)).to(torch.int32)
        feats = model.vit_model(packed_pixel_values=gi["packed_vit_tokens"],
                                packed_flattened_position_ids=gi["packed_vit_position_ids"], cu_seqlens=cu,
                                max_seqlen=int(gi["vit_token_seqlens"].max()))
        rc = model.forward_cache_update_vit(ns.qwen2_navit.NaiveCache(cfg.num_hidden_layers), **gi)
        tok = IntTokenizer()
        gt, kv2, rp2 = model.prepare_prompts(kv, rp, ["5 17 900", "8 8 100 4"], tok, NEW_TOKEN_IDS)
        rc = model.forward_cache_update_text(rc, **gt)
    with torch.no_grad():
        ofeats = osl.vit_forward(sd_full, vc, ogi["packed_vit_tokens"], ogi["packed_vit_position_ids"],
                                 ogi["vit_token_seqlens"])
        oc = osl.forward_cache_update_vit(sd_full, cfg, vc, om.KVCache(cfg.num_hidden_layers), **ogi)
        ogt, _, _ = obf.prepare_prompts(kv, rp, [[5, 17, 900], [8, 8, 100, 4]], 1000, 1001)
        oc = obf.forward_cache_update_text(sd_full, fc_of(cfg), oc, **ogt)
    assert torch.equal(feats, ofeats), "oracle ViT != reference"
    for li in range(cfg.num_hidden_layers):
        assert torch.equal(rc.key_cache[li], oc.key_cache[li]) and torch.equal(rc.value_cache[li], oc.value_cache[li])
    out["vit.features"] = feats.contiguous()
    out["vit.k_cache_last"] = rc.key_cache[cfg.num_hidden_layers - 1].contiguous()
    out["vit.v_cache_last"] = rc.value_cache[cfg.num_hidden_layers - 1].contiguous()
    print("ViT tower + ViT/text prefill: oracle == reference (bit-exact); feature |x| mean",
          float(feats.float().abs().mean()))
    save_file(out, os.path.join(OUT, "vit_tiny.safetensors"))


def golden_vae(ns):
    """FLUX VAE encode (with a fixed DiagonalGaussian noise) and decode, tiny 2-level config, 2 images 32x48."""
    from oracle import vae as ov
    ae_mod = ns.autoencoder
    params = ae_mod.AutoEncoderParams(resolution=32, in_channels=3, downsample=2, ch=128, out_ch=3, ch_mult=[1, 2],
                                      num_res_blocks=1, z_channels=16, scale_factor=0.3611, shift_factor=0.1159)
    ae = ae_mod.AutoEncoder(params).eval()
    sd = fixtures.vae_state_dict()
    ae.load_state_dict(sd, strict=True)
    vc = ov.VaeConfig(ch=128, ch_mult=[1, 2], num_res_blocks=1)
    img, noise = fixtures.vae_inputs()
    out = {}
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        moments = ae.encoder(img)
        ae.reg.sample = False
        z_mean = ae.encode(img)
        mean, logvar = torch.chunk(moments, 2, dim=1)
        z_noise = params.scale_factor * ((mean + torch.exp(0.5 * logvar) * noise) - params.shift_factor)
        z_in = torch.randn(2, 16, 16, 24, generator=torch.Generator().manual_seed(9))
        rec = ae.decode(z_in)
    with torch.no_grad():
        oz_mean = ov.encode(sd, vc, img)
        oz_noise = ov.encode(sd, vc, img, noise)
        orec = ov.decode(sd, vc, z_in)
    assert torch.equal(z_mean, oz_mean) and torch.equal(z_noise, oz_noise), "oracle VAE encode != reference"
    assert torch.equal(rec, orec), "oracle VAE decode != reference"
    out["vae.z_mean"], out["vae.z_noise"] = z_mean.contiguous(), z_noise.contiguous()
    out["vae.z_in"], out["vae.rec"] = z_in.contiguous(), rec.contiguous()
    # ---- VAE image as generation context (image-edit path): prepare_vae_images + forward_cache_update_vae ----
    cfg = fixtures.TINY_LM
    sdl = flow_state_dict(cfg, torch.bfloat16)
    lm_sd = {k[len("language_model."):]: v for k, v in sdl.items() if k.startswith("language_model.")}
    lm, rcfg = ref_lm(ns, cfg, lm_sd, torch.bfloat16)
    bcfg = ns.bagel.BagelConfig(visual_gen=True, visual_und=False, llm_config=rcfg, vit_config=None,
                                vae_config=SimpleNamespace(downsample=2, z_channels=16), latent_patch_size=2,
                                max_latent_size=16)
    model = ns.bagel.Bagel(lm, None, bcfg).eval()
    ref_shims.cast_parameters(model, torch.bfloat16)
    model.load_state_dict(sdl, strict=False)
    sd_full = dict(sdl)
    sd_full["latent_pos_embed.pos_embed"] = model.latent_pos_embed.pos_embed.data.clone()
    fc = obf.FlowConfig(lm=cfg, vae_downsample=2, max_latent_size=16)
    imgs = [img[0], img[1][:, :24, :32]]                       # ragged: 32x48 and 24x32
    gi, kv, rp = model.prepare_vae_images([0, 0], [0, 0], imgs, lambda im: im, NEW_TOKEN_IDS)
    ogi, okv, orp = obf.prepare_vae_images(fc, [0, 0], [0, 0], imgs, 1002, 1003)
    assert kv == okv and rp == orp
    for k in gi:
        same = (gi[k] == ogi[k]) if isinstance(gi[k], list) else torch.equal(gi[k], ogi[k])
        assert same, k
        if torch.is_tensor(gi[k]):
            out["vae_ctx." + k] = gi[k].clone()
    out["vae_ctx.kv_lens"], out["vae_ctx.ropes"] = torch.tensor(kv), torch.tensor(rp)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        rc = model.forward_cache_update_vae(ae, ns.qwen2_navit.NaiveCache(cfg.num_hidden_layers), **gi)
    with torch.no_grad():
        oc = obf.forward_cache_update_vae(sd_full, fc, lambda x: ov.encode(sd, vc, x), om.KVCache(cfg.num_hidden_layers), **ogi)
    for li in range(cfg.num_hidden_layers):
        assert torch.equal(rc.key_cache[li], oc.key_cache[li]) and torch.equal(rc.value_cache[li], oc.value_cache[li])
    out["vae_ctx.k_cache_last"] = rc.key_cache[cfg.num_hidden_layers - 1].contiguous()
    out["vae_ctx.v_cache_last"] = rc.value_cache[cfg.num_hidden_layers - 1].contiguous()
    print("VAE-context prefill (DiagonalGaussian sample=False): oracle == reference (bit-exact)")
    print("VAE encode/decode: oracle == reference (bit-exact); |rec| mean", float(rec.float().abs().mean()),
          "|z| mean", float(z_mean.float().abs().mean()))
    save_file(out, os.path.join(OUT, "vae_tiny.safetensors"))


def golden_inferencer(ns):
    """InterleaveInferencer end to end (reference inferencer.py:23-313) on a tiny LM + SigLIP tower + VAE:
    text->image, image+text->image (edit: VAE + ViT context, 3 CFG contexts), image+text->text (understanding),
    think->image (system prompt, generated text fed back as context). The reference's
    `torch.autocast(device_type="cuda")` block is inert on this CPU-only container, so the calls run under CPU autocast
    like every other fixture; DiagonalGaussian sampling is switched off (the noise draw is not reproducible)."""
    import numpy as np
    import warnings
    import data.transforms as rtf
    assert ns.inferencer is not None, getattr(ns, "inferencer_error", None)
    cfg, dtype, tv = fixtures.TINY_LM, torch.bfloat16, fixtures.TINY_VIT
    sd = flow_state_dict(cfg, dtype)
    sd.update(fixtures.vit_state_dict(tv["hidden"], tv["inter"], tv["layers"], tv["heads"], cfg.hidden_size, dtype=dtype))
    lm_sd = {k[len("language_model."):]: v for k, v in sd.items() if k.startswith("language_model.")}
    lm, rcfg = ref_lm(ns, cfg, lm_sd, dtype)
    sn = ns.siglip_navit
    vcfg = sn.SiglipVisionConfig(hidden_size=tv["hidden"], intermediate_size=tv["inter"], num_hidden_layers=tv["layers"],
                                 num_attention_heads=tv["heads"], num_channels=3, image_size=112, patch_size=14, rope=False)
    vit = sn.SiglipVisionModel(vcfg)
    vit.vision_model.embeddings.convert_conv2d_to_linear(vcfg)
    bcfg = ns.bagel.BagelConfig(visual_gen=True, visual_und=True, llm_config=rcfg, vit_config=vcfg,
                                vae_config=SimpleNamespace(downsample=2, z_channels=16), latent_patch_size=2,
                                max_latent_size=16, vit_max_num_patch_per_side=8)
    model = ns.bagel.Bagel(lm, vit, bcfg).eval()
    ref_shims.cast_parameters(model, dtype)
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all("pos_embed" in k for k in missing.missing_keys), missing
    ae_mod = ns.autoencoder
    params = ae_mod.AutoEncoderParams(resolution=32, in_channels=3, downsample=2, ch=128, out_ch=3, ch_mult=[1, 2],
                                      num_res_blocks=1, z_channels=16, scale_factor=0.3611, shift_factor=0.1159)
    ae = ae_mod.AutoEncoder(params).eval()
    ae.load_state_dict(fixtures.vae_state_dict(), strict=True)
    ae.reg.sample = False
    tok = fixtures.ToyTokenizer()
    inf = ns.inferencer.InterleaveInferencer(model, ae, tok, rtf.ImageTransform(64, 32, 4), rtf.ImageTransform(112, 56, 14),
                                             NEW_TOKEN_IDS)
    img = fixtures.inferencer_image()
    text = "5 17 900 33 2"
    kw = dict(num_timesteps=4, timestep_shift=3.0, cfg_text_scale=4.0, cfg_img_scale=1.5, cfg_interval=[0.4, 1.0],
              cfg_renorm_min=0.0, cfg_renorm_type="global")
    out = {"input.image": torch.from_numpy(np.asarray(img).copy())}

    def enc(s):
        return torch.tensor(list(s.encode("utf-8")), dtype=torch.uint8)

    def run(seed, **call):
        torch.manual_seed(seed)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
                r = inf(**call)
        # the reference leaves these set after an enable_taylorseer call; not used here, but keep the model clean
        return r

    r = run(21, text=text, image_shapes=(32, 48), **kw)
    out["t2i.image"] = torch.from_numpy(np.asarray(r["image"]).copy())
    r = run(22, image=img, text=text, **kw)
    out["edit.image"] = torch.from_numpy(np.asarray(r["image"]).copy())
    r = run(23, image=img, text=text, understanding_output=True, max_think_token_n=6, do_sample=False)
    out["und.text"] = enc(r["text"])
    r = run(24, text=text, think=True, max_think_token_n=5, do_sample=False, image_shapes=(32, 48), **kw)
    out["think.text"] = enc(r["text"])
    out["think.image"] = torch.from_numpy(np.asarray(r["image"]).copy())
    for k in ("t2i.image", "edit.image", "think.image"):
        print(f"inferencer[{k}]: shape {tuple(out[k].shape)}, mean level {out[k].float().mean():.1f}, "
              f"std {out[k].float().std():.1f}")
    print("inferencer[und.text]:", bytes(out["und.text"].tolist()).decode(), "| think.text:",
          bytes(out["think.text"].tolist()).decode())
    save_file(out, os.path.join(OUT, "inferencer_tiny.safetensors"))


def fc_of(cfg):
    return obf.FlowConfig(lm=cfg, max_latent_size=8)


def golden_lm_variants(ns):
    """The two other decoder-layer classes of Decoder_layer_dict (qwen2_navit.py:936-940): dense Qwen2DecoderLayer /
    PackedAttention (:236-378, 603-684) and Qwen2MoEDecoderLayer (:834-933). Causal und prefill (cache update), then a
    non-causal forward on top of the cache in mode "gen" (MoE routes the MLP; the dense model ignores the mode)."""
    out = {}
    for tag, cfg in (("dense", fixtures.TINY_DENSE_LM), ("moe", fixtures.TINY_MOE_LM)):
        dtype = torch.bfloat16
        sd = fixtures.lm_state_dict(cfg, seed=0, dtype=dtype)
        lm, _ = ref_lm(ns, cfg, sd, dtype)
        inp = fixtures.config1_inputs(cfg, dtype=dtype)
        n = 130
        xg = torch.randn(n, cfg.hidden_size, generator=torch.Generator().manual_seed(5)).to(dtype)
        kw = dict(query_lens=torch.tensor([n], dtype=torch.int32),
                  packed_query_position_ids=torch.full((n,), 512, dtype=torch.long),
                  packed_query_indexes=torch.arange(512, 512 + n), key_values_lens=torch.tensor([512], dtype=torch.int32),
                  packed_key_value_indexes=torch.arange(512), update_past_key_values=False, is_causal=False, mode="gen",
                  packed_vae_token_indexes=torch.arange(1, n - 1), packed_text_indexes=torch.tensor([0, n - 1]))
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            cache = ns.qwen2_navit.NaiveCache(cfg.num_hidden_layers)
            und = lm.forward_inference(
                packed_query_sequence=inp["x"], query_lens=inp["query_lens"],
                packed_query_position_ids=inp["und_position_ids"], packed_query_indexes=inp["query_indexes"],
                past_key_values=cache, key_values_lens=torch.tensor([0], dtype=torch.int32),
                packed_key_value_indexes=torch.zeros(0, dtype=torch.long), update_past_key_values=True, is_causal=True,
                mode="und")
            gen = lm.forward_inference(packed_query_sequence=xg, past_key_values=cache, **kw)
        with torch.no_grad():
            oc = om.KVCache(cfg.num_hidden_layers)
            oh, oc = om.lm_forward_inference(sd, cfg, inp["x"], inp["query_lens"], inp["und_position_ids"],
                                             inp["query_indexes"], oc, torch.tensor([0], dtype=torch.int32),
                                             torch.zeros(0, dtype=torch.long), True, True, "und")
            og, _ = om.lm_forward_inference(sd, cfg, xg, past_key_values=oc, **kw)
        assert torch.equal(oh, und.packed_query_sequence) and torch.equal(og, gen.packed_query_sequence), \
            f"oracle != reference ({tag})"
        for li in range(cfg.num_hidden_layers):
            assert torch.equal(oc.key_cache[li], cache.key_cache[li]) and torch.equal(oc.value_cache[li], cache.value_cache[li])
        out[tag + ".und_hidden"] = und.packed_query_sequence.contiguous()
        out[tag + ".gen_hidden"] = gen.packed_query_sequence.contiguous()
        out[tag + ".k_cache_last"] = cache.key_cache[cfg.num_hidden_layers - 1].contiguous()
        out[tag + ".v_cache_last"] = cache.value_cache[cfg.num_hidden_layers - 1].contiguous()
        print(f"lm_variants {tag} ({cfg.layer_module}): oracle == reference (bit-exact)")
    save_file(out, os.path.join(OUT, "lm_variants.safetensors"))


def golden_mode_b(ns):
    """dtype mode B — fp32 master weights under autocast, the way the eval drivers load the model
    (eval/gen/gen_images_mp.py:159-175, autocast :73): fp32 residual stream / norm outputs / RoPE tables.
    LM forward at head_dim 128 (lm_config1.safetensors already holds d64.B), and the text prefill + generate_image flow
    for both head dims."""
    out = {}
    dtype = torch.float32
    # ---- LM forward, d128 ----
    cfg = fixtures.TINY128_LM
    sd = fixtures.lm_state_dict(cfg, seed=0, dtype=dtype)
    lm, _ = ref_lm(ns, cfg, sd, dtype)
    inp = fixtures.config1_inputs(cfg, dtype=dtype)
    n = 130
    xg = torch.randn(n, cfg.hidden_size, generator=torch.Generator().manual_seed(5)).to(dtype)
    kw = dict(query_lens=torch.tensor([n], dtype=torch.int32), packed_query_position_ids=torch.full((n,), 512, dtype=torch.long),
              packed_query_indexes=torch.arange(512, 512 + n), key_values_lens=torch.tensor([512], dtype=torch.int32),
              packed_key_value_indexes=torch.arange(512), update_past_key_values=False, is_causal=False, mode="gen",
              packed_vae_token_indexes=torch.arange(1, n - 1), packed_text_indexes=torch.tensor([0, n - 1]))
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        cache = ns.qwen2_navit.NaiveCache(cfg.num_hidden_layers)
        und = lm.forward_inference(
            packed_query_sequence=inp["x"], query_lens=inp["query_lens"], packed_query_position_ids=inp["und_position_ids"],
            packed_query_indexes=inp["query_indexes"], past_key_values=cache, key_values_lens=torch.tensor([0], dtype=torch.int32),
            packed_key_value_indexes=torch.zeros(0, dtype=torch.long), update_past_key_values=True, is_causal=True, mode="und")
        gen = lm.forward_inference(packed_query_sequence=xg, past_key_values=cache, **kw)
    with torch.no_grad():
        oc = om.KVCache(cfg.num_hidden_layers)
        oh, oc = om.lm_forward_inference(sd, cfg, inp["x"], inp["query_lens"], inp["und_position_ids"], inp["query_indexes"],
                                         oc, torch.tensor([0], dtype=torch.int32), torch.zeros(0, dtype=torch.long), True, True, "und")
        og, _ = om.lm_forward_inference(sd, cfg, xg, past_key_values=oc, **kw)
    assert torch.equal(oh, und.packed_query_sequence) and torch.equal(og, gen.packed_query_sequence), "oracle != reference (d128 B)"
    assert und.packed_query_sequence.dtype == torch.float32
    out["d128.B.und_hidden"] = und.packed_query_sequence.contiguous()
    out["d128.B.gen_hidden"] = gen.packed_query_sequence.contiguous()
    out["d128.B.k_cache_last"] = cache.key_cache[cfg.num_hidden_layers - 1].contiguous()
    out["d128.B.v_cache_last"] = cache.value_cache[cfg.num_hidden_layers - 1].contiguous()
    print("mode B lm forward d128: oracle == reference (bit-exact)")
    # ---- flow ----
    tok = IntTokenizer()
    prompts = ["5 17 900 33 2", "8 8 100 4 77 650 12"]
    prompt_ids = [tok.encode(p) for p in prompts]
    image_sizes = [(64, 64), (64, 96)]
    for tag, cfg in (("d64", fixtures.TINY_LM), ("d128", fixtures.TINY128_LM)):
        sd = flow_state_dict(cfg, dtype)
        model = build_ref_bagel(ns, cfg, sd, dtype)
        sd_full = dict(sd)
        sd_full["latent_pos_embed.pos_embed"] = model.latent_pos_embed.pos_embed.data.clone()
        assert sd_full["latent_pos_embed.pos_embed"].dtype == torch.float32
        fc = obf.FlowConfig(lm=cfg, max_latent_size=8)

        def ctx(with_text):
            rc, oc_ = ns.qwen2_navit.NaiveCache(cfg.num_hidden_layers), om.KVCache(cfg.num_hidden_layers)
            kv, rp = [0, 0], [0, 0]
            if with_text:
                gi_, kv, rp = model.prepare_prompts(kv, rp, prompts, tok, NEW_TOKEN_IDS)
                ogi_, _, _ = obf.prepare_prompts([0, 0], [0, 0], prompt_ids, NEW_TOKEN_IDS["bos_token_id"], NEW_TOKEN_IDS["eos_token_id"])
                with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
                    rc = model.forward_cache_update_text(rc, **gi_)
                with torch.no_grad():
                    oc_ = obf.forward_cache_update_text(sd_full, fc, oc_, **ogi_)
                for li in range(cfg.num_hidden_layers):
                    assert torch.equal(rc.key_cache[li], oc_.key_cache[li])
            return rc, oc_, kv, rp

        rc_main, oc_main, kv_main, rp_main = ctx(True)
        rc_txt, oc_txt, kv_txt, rp_txt = ctx(False)
        rc_img, oc_img, kv_img, rp_img = ctx(True)
        out[f"{tag}.prefill.k_cache_last"] = rc_main.key_cache[cfg.num_hidden_layers - 1].contiguous()
        torch.manual_seed(2)
        gi = model.prepare_vae_latent(kv_main, rp_main, image_sizes, NEW_TOKEN_IDS)
        cfg_t = model.prepare_vae_latent_cfg(kv_txt, rp_txt, image_sizes)
        cfg_i = model.prepare_vae_latent_cfg(kv_img, rp_img, image_sizes)

        def obranch(d, cache):
            return dict(packed_position_ids=d["cfg_packed_position_ids"], packed_query_indexes=d["cfg_packed_query_indexes"],
                        key_values_lens=d["cfg_key_values_lens"], past_key_values=cache,
                        packed_key_value_indexes=d["cfg_packed_key_value_indexes"])

        for name, sT, sI, rt in [("nocfg", 1.0, 1.0, "global"), ("global_img", 4.0, 1.5, "global"),
                                 ("text_channel_img", 4.0, 1.5, "text_channel")]:
            kwargs = dict(num_timesteps=4, timestep_shift=3.0, cfg_renorm_min=0.0, cfg_renorm_type=rt, cfg_interval=[0.4, 1.0],
                          cfg_text_scale=sT, cfg_img_scale=sI)
            ref_kw = dict(kwargs)
            ref_kw.update(
                cfg_text_packed_position_ids=cfg_t["cfg_packed_position_ids"], cfg_text_packed_query_indexes=cfg_t["cfg_packed_query_indexes"],
                cfg_text_key_values_lens=cfg_t["cfg_key_values_lens"], cfg_text_packed_key_value_indexes=cfg_t["cfg_packed_key_value_indexes"],
                cfg_text_past_key_values=rc_txt,
                cfg_img_packed_position_ids=cfg_i["cfg_packed_position_ids"], cfg_img_packed_query_indexes=cfg_i["cfg_packed_query_indexes"],
                cfg_img_key_values_lens=cfg_i["cfg_key_values_lens"], cfg_img_packed_key_value_indexes=cfg_i["cfg_packed_key_value_indexes"],
                cfg_img_past_key_values=rc_img)
            with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
                lat = model.generate_image(past_key_values=rc_main, **gi, **ref_kw)
            with torch.no_grad():
                olat = obf.generate_image(sd_full, fc, dict(gi), oc_main, cfg_text=obranch(cfg_t, oc_txt),
                                          cfg_img=obranch(cfg_i, oc_img), **kwargs)
            for a, b in zip(lat, olat):
                assert torch.equal(a, b), f"oracle generate_image != reference (mode B {tag} {name})"
            out[f"{tag}.gen.{name}.latents"] = torch.cat(lat, dim=0).contiguous()
            print(f"mode B generate_image[{tag} {name}]: oracle == reference (bit-exact)")
    save_file(out, os.path.join(OUT, "mode_b_tiny.safetensors"))


def golden_train_forward(ns):
    """Bagel.forward in training mode (bagel.py:101-229; forward_train paths qwen2_navit.py:406-497, 713-755, 970-1016)
    with dense per-sample masks (data/data_utils.py:72-103): losses + last hidden state of the reference; also asserts
    that oracle/train_forward.py reproduces it bit for bit."""
    from oracle import siglip as osl, train_forward as otf
    model, sd_full, cfg, tv = _ref_bagel_with_vit(ns, rope=False)
    model.config.timestep_shift = 1.0
    b = fixtures.train_batch()
    masks = [ns.data_utils.prepare_attention_mask_per_sample(s, m) for s, m in zip(b["nested_split_lens"], b["nested_attn_modes"])]
    omasks = [otf.prepare_attention_mask_per_sample(s, m) for s, m in zip(b["nested_split_lens"], b["nested_attn_modes"])]
    for a, c in zip(masks, omasks):
        assert torch.equal(a, c)
    hidden = {}
    hook = model.language_model.model.register_forward_hook(lambda m, i, o: hidden.__setitem__("h", o.detach().clone()))
    model.train()
    kw = {k: v for k, v in b.items() if k not in ("split_lens", "attn_modes", "nested_split_lens", "nested_attn_modes")}
    torch.manual_seed(7)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        out = model(nested_attention_masks=masks, **kw)
    hook.remove()
    model.eval()
    # the same noise draw the reference made (first RNG use inside forward: torch.randn_like(packed_latent_clean), :184)
    n_lat = sum(h * w for h, w in b["patchified_vae_latent_shapes"])
    torch.manual_seed(7)
    noise = torch.randn(n_lat, 64)
    vc = osl.VitConfig(hidden_size=tv["hidden"], intermediate_size=tv["inter"], num_hidden_layers=tv["layers"],
                       num_attention_heads=tv["heads"])
    fc = obf.FlowConfig(lm=cfg, max_latent_size=8)
    with torch.no_grad():
        o = otf.bagel_forward_train(
            sd_full, fc, b["sequence_length"], b["packed_text_ids"], b["packed_text_indexes"], b["sample_lens"],
            b["packed_position_ids"], omasks, noise, timestep_shift=1.0, ce_loss_indexes=b["ce_loss_indexes"],
            packed_label_ids=b["packed_label_ids"],
            vit=(vc, b["packed_vit_tokens"], b["packed_vit_token_indexes"], b["packed_vit_position_ids"], b["vit_token_seqlens"]),
            padded_latent=b["padded_latent"], patchified_vae_latent_shapes=b["patchified_vae_latent_shapes"],
            packed_latent_position_ids=b["packed_latent_position_ids"], packed_vae_token_indexes=b["packed_vae_token_indexes"],
            packed_timesteps=b["packed_timesteps"], mse_loss_indexes=b["mse_loss_indexes"])
    assert torch.equal(o["last_hidden_state"], hidden["h"]), "oracle forward_train hidden != reference"
    assert torch.equal(o["mse"], out["mse"]) and torch.equal(o["ce"], out["ce"]), "oracle losses != reference"
    print("train forward: oracle == reference (bit-exact); mse mean", float(out["mse"].mean()), "ce mean", float(out["ce"].mean()))
    save_file({"train.mse": out["mse"].contiguous(), "train.ce": out["ce"].contiguous(),
               "train.last_hidden_state": hidden["h"].contiguous(), "train.noise": noise},
              os.path.join(OUT, "train_forward_tiny.safetensors"))


def _ref_bagel_with_vit(ns, rope):
    cfg = fixtures.TINY_LM
    dtype = torch.bfloat16
    tv = fixtures.TINY_VIT
    sd = flow_state_dict(cfg, dtype)
    sd.update(fixtures.vit_state_dict(tv["hidden"], tv["inter"], tv["layers"], tv["heads"], cfg.hidden_size, dtype=dtype))
    lm_sd = {k[len("language_model."):]: v for k, v in sd.items() if k.startswith("language_model.")}
    lm, rcfg = ref_lm(ns, cfg, lm_sd, dtype)
    sn = ns.siglip_navit
    vcfg = sn.SiglipVisionConfig(hidden_size=tv["hidden"], intermediate_size=tv["inter"], num_hidden_layers=tv["layers"],
                                 num_attention_heads=tv["heads"], num_channels=3, image_size=112, patch_size=14, rope=rope)
    vit = sn.SiglipVisionModel(vcfg)
    vit.vision_model.embeddings.convert_conv2d_to_linear(vcfg)
    bcfg = ns.bagel.BagelConfig(visual_gen=True, visual_und=True, llm_config=rcfg, vit_config=vcfg,
                                vae_config=SimpleNamespace(downsample=8, z_channels=16), latent_patch_size=2,
                                max_latent_size=8, vit_max_num_patch_per_side=8)
    model = ns.bagel.Bagel(lm, vit, bcfg).eval()
    ref_shims.cast_parameters(model, dtype)
    missing = model.load_state_dict(sd, strict=False)
    # rope=True: no learned position table in the tower (its key is then unexpected), RoPE tables are buffers
    assert all("position_embedding" in k for k in missing.unexpected_keys) and (rope or not missing.unexpected_keys), missing
    assert all("pos_embed" in k or ".rope." in k for k in missing.missing_keys), missing
    sd_full = dict(sd)
    sd_full["latent_pos_embed.pos_embed"] = model.latent_pos_embed.pos_embed.data.clone()
    sd_full["vit_pos_embed.pos_embed"] = model.vit_pos_embed.pos_embed.data.clone()
    return model, sd_full, cfg, tv


def golden_vit_rope(ns):
    """SigLIP tower with the optional 2-D RoPE enabled (siglip_navit.py:102-142, 224-230, 343-365): same weights and
    images as vit_tiny.safetensors, config.rope=True, head_dim 72 -> 36-wide row / column halves."""
    from oracle import siglip as osl
    model, sd_full, cfg, tv = _ref_bagel_with_vit(ns, rope=True)
    vc = osl.VitConfig(hidden_size=tv["hidden"], intermediate_size=tv["inter"], num_hidden_layers=tv["layers"],
                       num_attention_heads=tv["heads"], rope=True, image_size=112)
    images = fixtures.vit_images()
    gi, kv, rp = model.prepare_vit_images([0, 0], [0, 0], images, lambda im: im, NEW_TOKEN_IDS)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        cu = torch.nn.functional.pad(torch.cumsum(gi["vit_token_seqlens"], 0), (1, 0)).to(torch.int32)
        feats = model.vit_model(packed_pixel_values=gi["packed_vit_tokens"],
                                packed_flattened_position_ids=gi["packed_vit_position_ids"], cu_seqlens=cu,
                                max_seqlen=int(gi["vit_token_seqlens"].max()))
    with torch.no_grad():
        ofeats = osl.vit_forward(sd_full, vc, gi["packed_vit_tokens"], gi["packed_vit_position_ids"], gi["vit_token_seqlens"])
    assert torch.equal(feats, ofeats), "oracle ViT (rope=True) != reference"
    print("ViT tower rope=True: oracle == reference (bit-exact); feature |x| mean", float(feats.float().abs().mean()))
    save_file({"vit_rope.features": feats.contiguous()}, os.path.join(OUT, "vit_rope_tiny.safetensors"))


def golden_chat(ns):
    """Bagel.chat (bagel.py:1004-1075): two images + a prompt -> greedy text, through the reference's own method."""
    model, sd_full, cfg, tv = _ref_bagel_with_vit(ns, rope=False)
    tok = fixtures.ToyTokenizer()
    images = fixtures.vit_images()
    import warnings
    logits = []
    hook = model.language_model.lm_head.register_forward_hook(lambda m, i, o: logits.append(o.detach().clone()))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            text = model.chat(tok, dict(NEW_TOKEN_IDS), lambda im: im, images, "5 17 900 33 2 describe", max_length=8,
                              do_sample=False)
    hook.remove()
    lg = torch.stack([x.reshape(-1) for x in logits], 0)          # [steps, V] (one sample)
    top2 = lg.float().topk(2, dim=-1).values
    print("chat:", repr(text), "| top-1/top-2 logit margins per step:", [round(float(x), 3) for x in top2[:, 0] - top2[:, 1]])
    save_file({"chat.text": torch.tensor(list(text.encode("utf-8")), dtype=torch.uint8), "chat.logits": lg.contiguous()},
              os.path.join(OUT, "chat_tiny.safetensors"))


def main():
    if not ref_shims.reference_available():
        raise SystemExit("reference tree not available; fixtures can only be generated in the build container")
    torch.set_num_threads(8)
    ns = ref_shims.load_reference()
    if "--new-only" in sys.argv:    # round-2 additions only (leaves the round-1 fixture files untouched)
        golden_lm_variants(ns)
        golden_vit_rope(ns)
        golden_chat(ns)
        golden_mode_b(ns)
        golden_train_forward(ns)
        return
    golden_lm_config1(ns)
    golden_flow(ns)
    golden_vit(ns)
    golden_vae(ns)
    golden_inferencer(ns)
    golden_lm_variants(ns)
    golden_vit_rope(ns)
    golden_chat(ns)
    golden_mode_b(ns)
    golden_train_forward(ns)
    sizes = {f: os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT) if f.endswith(".safetensors")}
    print("wrote", sizes)


if __name__ == "__main__":
    main()
