"""GPU diagnostic: product LM forward + generate_image vs the committed golden fixtures (reference outputs)."""
import sys, time
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from safetensors.torch import load_file
from oracle import fixtures
import helpers
from bagel_b200.qwen2_navit import NaiveCache

def rep(name, out, ref):
    out = out.float().cpu(); ref = ref.float().cpu()
    err = (out - ref).abs()
    rel = err.max() / ref.abs().max()
    print(f"[{name}] max_abs_err={err.max().item():.4e} mean_abs_err={err.mean().item():.3e} ref_absmax={ref.abs().max().item():.3f} "
          f"rel_to_max={rel.item():.3e} finite={bool(torch.isfinite(out).all())}", flush=True)

g = load_file("tests/golden/lm_config1.safetensors")
for tag, cfg in (("d64", fixtures.TINY_LM), ("d128", fixtures.TINY128_LM)):
    model = helpers.build_product_bagel(cfg, "cuda")
    lm = model.language_model
    inp = fixtures.config1_inputs(cfg)
    cache = NaiveCache(cfg.num_hidden_layers)
    und = lm.forward_inference(packed_query_sequence=inp["x"], query_lens=inp["query_lens"], packed_query_position_ids=inp["und_position_ids"],
        packed_query_indexes=inp["query_indexes"], past_key_values=cache, key_values_lens=torch.tensor([0], dtype=torch.int32),
        packed_key_value_indexes=torch.zeros(0, dtype=torch.long), update_past_key_values=True, is_causal=True, mode="und")
    torch.cuda.synchronize()
    rep(f"{tag} und hidden", und.packed_query_sequence, g[f"{tag}.A.und_hidden"])
    rep(f"{tag} k_cache_last", cache.key_cache[cfg.num_hidden_layers - 1], g[f"{tag}.A.k_cache_last"])
    rep(f"{tag} v_cache_last", cache.value_cache[cfg.num_hidden_layers - 1], g[f"{tag}.A.v_cache_last"])
    n = 130
    xg = torch.randn(n, cfg.hidden_size, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16)
    gen = lm.forward_inference(packed_query_sequence=xg, query_lens=torch.tensor([n], dtype=torch.int32),
        packed_query_position_ids=torch.full((n,), 512, dtype=torch.long), packed_query_indexes=torch.arange(512, 512 + n),
        key_values_lens=torch.tensor([512], dtype=torch.int32), packed_key_value_indexes=torch.arange(512), update_past_key_values=False,
        is_causal=False, mode="gen", packed_vae_token_indexes=torch.arange(1, n - 1), packed_text_indexes=torch.tensor([0, n - 1]),
        past_key_values=cache)
    torch.cuda.synchronize()
    rep(f"{tag} gen hidden", gen.packed_query_sequence, g[f"{tag}.A.gen_hidden"])

# ---- flow ----
gf = load_file("tests/golden/flow_tiny.safetensors")
cfg = fixtures.TINY_LM
model = helpers.build_product_bagel(cfg, "cuda")
tok = helpers.IntTokenizer()
def ctx(with_text):
    c = NaiveCache(cfg.num_hidden_layers); kv, rp = [0, 0], [0, 0]
    if with_text:
        gi, kv, rp = model.prepare_prompts(kv, rp, helpers.PROMPTS, tok, helpers.NEW_TOKEN_IDS)
        c = model.forward_cache_update_text(c, **gi)
    return c, kv, rp
c_main, kv_main, rp_main = ctx(True); c_txt, kv_txt, rp_txt = ctx(False); c_img, kv_img, rp_img = ctx(True)
rep("prefill k_cache_last", c_main.key_cache[cfg.num_hidden_layers - 1], gf["prefill.k_cache_last"])
torch.manual_seed(2)
gi = model.prepare_vae_latent(kv_main, rp_main, helpers.IMAGE_SIZES, helpers.NEW_TOKEN_IDS)
for k in gi:
    assert torch.equal(gi[k], gf["latent." + k]), k
ct = model.prepare_vae_latent_cfg(kv_txt, rp_txt, helpers.IMAGE_SIZES); ci = model.prepare_vae_latent_cfg(kv_img, rp_img, helpers.IMAGE_SIZES)
for name, sT, sI, rt in [("nocfg", 1.0, 1.0, "global"), ("global", 4.0, 1.0, "global"), ("channel", 4.0, 1.0, "channel"),
                         ("global_img", 4.0, 1.5, "global"), ("text_channel_img", 4.0, 1.5, "text_channel")]:
    lat = model.generate_image(past_key_values=c_main, **gi, num_timesteps=4, timestep_shift=3.0, cfg_renorm_min=0.0, cfg_renorm_type=rt,
        cfg_interval=[0.4, 1.0], cfg_text_scale=sT, cfg_img_scale=sI,
        cfg_text_packed_position_ids=ct["cfg_packed_position_ids"], cfg_text_packed_query_indexes=ct["cfg_packed_query_indexes"],
        cfg_text_key_values_lens=ct["cfg_key_values_lens"], cfg_text_packed_key_value_indexes=ct["cfg_packed_key_value_indexes"],
        cfg_text_past_key_values=c_txt,
        cfg_img_packed_position_ids=ci["cfg_packed_position_ids"], cfg_img_packed_query_indexes=ci["cfg_packed_query_indexes"],
        cfg_img_key_values_lens=ci["cfg_key_values_lens"], cfg_img_packed_key_value_indexes=ci["cfg_packed_key_value_indexes"],
        cfg_img_past_key_values=c_img)
    torch.cuda.synchronize()
    rep(f"generate_image[{name}]", torch.cat(lat, 0), gf[f"gen.{name}.latents"])
    # the velocity part of the update: x_T - x_0
    rep(f"   delta[{name}]", torch.cat(lat, 0).cpu() - gi["packed_init_noises"], gf[f"gen.{name}.latents"] - gi["packed_init_noises"])
print("E2E_DONE")
