"""-m gpu: the parts of the reference's model surface added in round 2, each against outputs of the UNMODIFIED reference
(tests/golden/make_golden.py --new-only): the dense `Qwen2DecoderLayer` / `PackedAttention` route and the
`Qwen2MoEDecoderLayer` (qwen2_navit.py:236-378, 603-684, 834-933), the SigLIP tower with 2-D RoPE
(siglip_navit.py:102-142, 224-230), `Bagel.chat` (bagel.py:1004-1075) and the CUDA-graph / workspace interplay."""
import os

import pytest
import torch
from safetensors.torch import load_file

import helpers
from oracle import fixtures, qwen2_mot as om
from test_gpu_model import _check, _f32

pytestmark = pytest.mark.gpu


def _product_lm(cfg, device="cuda"):
    from bagel_b200.config import Qwen2Config
    from bagel_b200.qwen2_navit import Qwen2ForCausalLM
    llm = Qwen2Config(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                      num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                      num_key_value_heads=cfg.num_key_value_heads, rope_theta=cfg.rope_theta,
                      rms_norm_eps=cfg.rms_norm_eps, qk_norm=True, layer_module=cfg.layer_module)
    lm = Qwen2ForCausalLM(llm, device=device)
    lm.load_state_dict(fixtures.lm_state_dict(cfg, seed=0))
    return lm


@pytest.mark.parametrize("tag,cfg", [("dense", fixtures.TINY_DENSE_LM), ("moe", fixtures.TINY_MOE_LM)])
def test_dense_and_moe_decoder_layers(golden_dir, tag, cfg):
    """Causal und prefill with cache update, then a non-causal mode="gen" forward on top of the cache: the dense model
    ignores the mode (one expert, PackedAttention bf16 flow); the MoE model shares attention + norms and routes only
    the MLP and the final norm (text rows -> mlp / norm, latent rows -> mlp_moe_gen / norm_moe_gen)."""
    from bagel_b200.qwen2_navit import NaiveCache
    g = load_file(os.path.join(golden_dir, "lm_variants.safetensors"))
    lm = _product_lm(cfg)
    assert lm.model.layer_kind == tag and lm.model.use_moe == (tag == "moe")
    inp = fixtures.config1_inputs(cfg)
    cache = NaiveCache(cfg.num_hidden_layers)
    kw_und = dict(query_lens=inp["query_lens"], packed_query_position_ids=inp["und_position_ids"],
                  packed_query_indexes=inp["query_indexes"], key_values_lens=torch.tensor([0], dtype=torch.int32),
                  packed_key_value_indexes=torch.zeros(0, dtype=torch.long), update_past_key_values=True,
                  is_causal=True, mode="und")
    und = lm.forward_inference(packed_query_sequence=inp["x"], past_key_values=cache, **kw_und)
    n = 130
    xg = torch.randn(n, cfg.hidden_size, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16)
    kw_gen = dict(query_lens=torch.tensor([n], dtype=torch.int32),
                  packed_query_position_ids=torch.full((n,), 512, dtype=torch.long),
                  packed_query_indexes=torch.arange(512, 512 + n), key_values_lens=torch.tensor([512], dtype=torch.int32),
                  packed_key_value_indexes=torch.arange(512), update_past_key_values=False, is_causal=False, mode="gen",
                  packed_vae_token_indexes=torch.arange(1, n - 1), packed_text_indexes=torch.tensor([0, n - 1]))
    gen = lm.forward_inference(packed_query_sequence=xg, past_key_values=cache, **kw_gen)
    torch.cuda.synchronize()
    sd32 = _f32(fixtures.lm_state_dict(cfg, seed=0))
    with torch.no_grad(), om.high_precision():
        oc = om.KVCache(cfg.num_hidden_layers)
        t_und, oc = om.lm_forward_inference(sd32, cfg, inp["x"].float(), past_key_values=oc, **kw_und)
        t_gen, _ = om.lm_forward_inference(sd32, cfg, xg.float(), past_key_values=oc, **kw_gen)
    last = cfg.num_hidden_layers - 1
    _check("und hidden", und.packed_query_sequence, g[tag + ".und_hidden"], t_und)
    _check("k cache", cache.key_cache[last], g[tag + ".k_cache_last"], oc.key_cache[last])
    _check("v cache", cache.value_cache[last], g[tag + ".v_cache_last"], oc.value_cache[last])
    _check("gen hidden", gen.packed_query_sequence, g[tag + ".gen_hidden"], t_gen)


def test_unknown_layer_module_and_missing_heads_are_rejected():
    from bagel_b200.config import Qwen2Config
    from bagel_b200.qwen2_navit import Qwen2ForCausalLM
    with pytest.raises(ValueError):
        Qwen2ForCausalLM(Qwen2Config(hidden_size=256, num_attention_heads=4, num_key_value_heads=2, intermediate_size=512,
                                     num_hidden_layers=1, vocab_size=64, layer_module="Qwen2MoXDecoderLayer"), device="cuda")
    model = helpers.build_product_bagel(fixtures.TINY_LM, "cuda", load=False)
    sd = helpers.flow_state_dict(fixtures.TINY_LM)
    sd.pop("llm2vae.weight")
    with pytest.raises(KeyError, match="llm2vae.weight"):
        model.load_state_dict(sd)
    sd = helpers.flow_state_dict(fixtures.TINY_LM)
    sd["surprise.weight"] = torch.zeros(1)
    model.load_state_dict(sd)                      # strict=False: extras ignored
    with pytest.raises(KeyError, match="surprise"):
        model.load_state_dict(sd, strict=True)


def test_siglip_rope2d_tower(golden_dir):
    """SigLIP NaViT with config.rope=True: no learned position table; q/k halves rotated by the row / column tables."""
    from bagel_b200.config import SiglipVisionConfig
    from bagel_b200.siglip_navit import SiglipVisionModel
    from oracle import siglip as osl
    g = load_file(os.path.join(golden_dir, "vit_rope_tiny.safetensors"))
    gin = load_file(os.path.join(golden_dir, "vit_tiny.safetensors"))
    tv = fixtures.TINY_VIT
    cfg = fixtures.TINY_LM
    vcfg = SiglipVisionConfig(hidden_size=tv["hidden"], intermediate_size=tv["inter"], num_hidden_layers=tv["layers"],
                              num_attention_heads=tv["heads"], num_channels=3, image_size=112, patch_size=14, rope=True)
    vit = SiglipVisionModel(vcfg, device="cuda")
    sd = helpers.vit_flow_state_dict(cfg)
    vit.load_state_dict({k[len("vit_model."):]: v for k, v in sd.items() if k.startswith("vit_model.")})
    vl = gin["vit_in.vit_token_seqlens"]
    cu = torch.cat([torch.zeros(1, dtype=torch.int64), vl.to(torch.int64).cumsum(0)]).to(torch.int32)
    feats = vit(packed_pixel_values=gin["vit_in.packed_vit_tokens"],
                packed_flattened_position_ids=gin["vit_in.packed_vit_position_ids"], cu_seqlens=cu, max_seqlen=int(vl.max()))
    vc = osl.VitConfig(hidden_size=tv["hidden"], intermediate_size=tv["inter"], num_hidden_layers=tv["layers"],
                       num_attention_heads=tv["heads"], rope=True, image_size=112)
    with torch.no_grad(), om.high_precision():
        truth = osl.vit_forward(_f32(sd), vc, gin["vit_in.packed_vit_tokens"], gin["vit_in.packed_vit_position_ids"], vl)
    _check("vit rope features", feats, g["vit_rope.features"], truth, max_ulps_of_scale=8.0)
    # and it is a different function from the rope=False tower
    assert (feats.float().cpu() - gin["vit.features"].float()).abs().max() > 0.1


def test_rope2d_kernel_bit_exact_vs_torch():
    """bagel_siglip_rope2d_bf16 against the reference's own expression (fp32 tables x bf16 q => fp32 math, one rounding)."""
    from bagel_b200 import ops
    from bagel_b200.siglip_navit import _rope2d_tables
    g = torch.Generator(device="cuda").manual_seed(3)
    n, heads, d, stride, side = 77, 6, 72, 128, 8
    x = torch.randn(n, heads * stride + 128, device="cuda", generator=g).to(torch.bfloat16)
    pos = torch.randint(0, side * side, (n,), device="cuda", generator=g)
    tabs = [t.cuda().contiguous() for t in _rope2d_tables(d // 2, side, side)]
    ref = x.clone()
    xv = x[:, : heads * stride].view(n, heads, stride)[..., :d]

    def rot(v):
        h = v.shape[-1] // 2
        return torch.cat((-v[..., h:], v[..., :h]), dim=-1)

    ch, sh, cw, sw = (t[pos].unsqueeze(1) for t in tabs)
    a, b = xv[..., : d // 2], xv[..., d // 2:]
    want = torch.cat([a * ch + rot(a) * sh, b * cw + rot(b) * sw], dim=-1).to(torch.bfloat16)
    ops.siglip_rope2d(x, heads, stride, d, pos, *tabs)
    got = x[:, : heads * stride].view(n, heads, stride)
    assert torch.equal(got[..., :d], want)
    assert torch.equal(got[..., d:], ref[:, : heads * stride].view(n, heads, stride)[..., d:])   # padding untouched
    assert torch.equal(x[:, heads * stride:], ref[:, heads * stride:])


def test_chat_matches_reference(golden_dir):
    """Bagel.chat: two images -> SigLIP prefill each, prompt prefill, greedy decode; the reference's own chat() output.
    (Every decision of the reference had a top-1/top-2 margin >= 0.15, twice the bf16 noise of these logits.)"""
    g = load_file(os.path.join(golden_dir, "chat_tiny.safetensors"))
    want = bytes(g["chat.text"].tolist()).decode("utf-8")
    model = helpers.build_product_bagel_with_vit(fixtures.TINY_LM, "cuda")
    tok = fixtures.ToyTokenizer()
    got = model.chat(tok, dict(helpers.NEW_TOKEN_IDS), lambda im: im, fixtures.vit_images(), "5 17 900 33 2 describe",
                     max_length=8, do_sample=False)
    assert got == want, (got, want)


def test_stale_graph_is_dropped_when_a_workspace_is_reallocated():
    """A captured step graph holds raw pointers into the LM's grow-only workspaces. Run a LARGER forward between two
    steps of a planned run (it re-allocates them): the runner must notice (generation counter) and re-capture instead
    of replaying into freed memory; results equal an uninterrupted run bit for bit."""
    from bagel_b200.qwen2_navit import NaiveCache
    cfg = fixtures.TINY_LM
    model = helpers.build_product_bagel(cfg, "cuda")
    tok = helpers.IntTokenizer()

    def ctx(with_text):
        c, kv, rp = NaiveCache(cfg.num_hidden_layers), [0, 0], [0, 0]
        if with_text:
            gi_, kv, rp = model.prepare_prompts(kv, rp, helpers.PROMPTS, tok, helpers.NEW_TOKEN_IDS)
            c = model.forward_cache_update_text(c, **gi_)
        return c, kv, rp

    def make():
        c_main, kv_m, rp_m = ctx(True)
        c_txt, kv_t, rp_t = ctx(False)
        torch.manual_seed(2)
        gi = model.prepare_vae_latent(kv_m, rp_m, helpers.IMAGE_SIZES, helpers.NEW_TOKEN_IDS)
        ct = model.prepare_vae_latent_cfg(kv_t, rp_t, helpers.IMAGE_SIZES)
        return model.make_flow_runner(
            past_key_values=c_main, **gi, num_timesteps=9, timestep_shift=3.0, cfg_renorm_type="global",
            cfg_interval=[0.3, 0.8], cfg_text_scale=3.0, cfg_text_packed_position_ids=ct["cfg_packed_position_ids"],
            cfg_text_packed_query_indexes=ct["cfg_packed_query_indexes"], cfg_text_key_values_lens=ct["cfg_key_values_lens"],
            cfg_text_packed_key_value_indexes=ct["cfg_packed_key_value_indexes"], cfg_text_past_key_values=c_txt)

    model.use_cuda_graph = True
    r = make()
    assert r.use_cuda_graph
    for i in range(r.num_steps):
        r.step(i)
    want = [x.clone() for x in r.latents()]
    r = make()
    big = torch.randn(4000, cfg.hidden_size, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
    for i in range(r.num_steps):
        r.step(i)
        if i == 3:   # graphs for both branch sets exist by now; this forward needs 4000-row workspaces
            gen0 = model.language_model.model._ws_gen
            model.language_model.forward_inference(
                packed_query_sequence=big, query_lens=torch.tensor([4000], dtype=torch.int32),
                packed_query_position_ids=torch.arange(4000), packed_query_indexes=torch.arange(4000),
                past_key_values=NaiveCache(cfg.num_hidden_layers), key_values_lens=torch.tensor([0], dtype=torch.int32),
                packed_key_value_indexes=torch.zeros(0, dtype=torch.long), update_past_key_values=False, is_causal=True,
                mode="und")
            assert model.language_model.model._ws_gen > gen0
    torch.cuda.synchronize()
    got = r.latents()
    assert all(torch.equal(a, b) for a, b in zip(got, want))
