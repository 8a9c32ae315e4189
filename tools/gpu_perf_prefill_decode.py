"""BASELINE configs[2] (image-understanding prefill: 32 x (SigLIP 378^2 = 729 patches + 512 text tokens)) and the text
decode loop at BAGEL-7B shapes, random-init weights. Prints prefill tokens/s and decode tokens/s."""
import sys, time, torch
sys.path.insert(0, ".")
from bagel_b200 import synthetic
from bagel_b200.config import SiglipVisionConfig
from bagel_b200.modeling_utils import MLPconnector, PositionEmbedding
from bagel_b200.qwen2_navit import NaiveCache
from bagel_b200.siglip_navit import SiglipVisionModel
from oracle import fixtures

dev = "cuda"
model = synthetic.build_random_bagel(device=dev, seed=0)
L = model.config.llm_config.num_hidden_layers
# attach a random SigLIP-so400m tower + connector
vcfg = SiglipVisionConfig(hidden_size=1152, intermediate_size=4304, num_hidden_layers=26, num_attention_heads=16,
                          num_channels=3, image_size=980, patch_size=14, rope=False)
vsd = fixtures.vit_state_dict(1152, 4304, 26, 16, 3584, max_side=70, seed=5)
vit = SiglipVisionModel(vcfg, dev); vit.load_state_dict({k[len("vit_model."):]: v for k, v in vsd.items() if k.startswith("vit_model.")})
model.vit_model = vit; model.config.vit_config = vcfg; model.config.visual_und = True
model.vit_patch_size, model.vit_max_num_patch_per_side, model.vit_hidden_size = 14, 70, 1152
model.connector = MLPconnector(1152, 3584); model.connector.load(vsd, "connector.", dev)
model.vit_pos_embed = PositionEmbedding(70, 3584, dev)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
imgs = [torch.rand(3, 378, 378, generator=torch.Generator().manual_seed(3 + i)) * 2 - 1 for i in range(B)]
tok = synthetic.RandomIdTokenizer(1)

def prefill():
    cache = NaiveCache(L)
    gi, kv, rp = model.prepare_vit_images([0] * B, [0] * B, imgs, lambda im: im, synthetic.NEW_TOKEN_IDS)
    cache = model.forward_cache_update_vit(cache, **gi)
    gt, kv, rp = model.prepare_prompts(kv, rp, ["512"] * B, tok, synthetic.NEW_TOKEN_IDS)
    cache = model.forward_cache_update_text(cache, **gt)
    return cache, kv, rp

prefill(); torch.cuda.synchronize()
t0 = time.perf_counter(); cache, kv, rp = prefill(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
ntok = sum(kv)
print(f"[cfg3 und prefill] B={B}: {ntok} tokens ({ntok//B}/sample) in {dt*1e3:.1f} ms = {ntok/dt:.0f} tokens/s (incl. host packing, H2D of patches)", flush=True)

gs = model.prepare_start_tokens(kv, rp, synthetic.NEW_TOKEN_IDS)
times = {}
for steps in (4, 32, 160):
    from copy import deepcopy
    c = deepcopy(cache); torch.cuda.synchronize()
    t0 = time.perf_counter(); toks = model.generate_text(past_key_values=c, max_length=steps, do_sample=False, **gs); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"[decode] B={B} ctx={ntok//B}: {steps} steps in {dt*1e3:.1f} ms = {dt/steps*1e3:.2f} ms/step = {B*steps/dt:.0f} tokens/s "
          f"(HBM roofline: ~14.1 GB weights/step -> 2.1 ms)", flush=True)
    times[steps] = dt
marg = (times[160] - times[32]) / 128
print(f"[decode] steady state (marginal over steps 32..160, CUDA-graph replay; a negative or odd value means a run hit an outlier): {marg*1e3:.2f} ms/step = {B/marg:.0f} tokens/s", flush=True)
