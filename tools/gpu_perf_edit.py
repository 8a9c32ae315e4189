"""BASELINE configs[3] on one GPU: image edit through the public orchestrator — VAE encode of a 1024^2 image + SigLIP
context (980^2 -> 4900 patches) + 64-token prompt, then a 50-timestep denoise with THREE CFG branches (text scale 4,
image scale 2) and the VAE decode. BAGEL-7B shapes, random-init weights, one sample per call (the reference's
inferencer is single-sample; the 8-GPU configuration shards 16 such samples, 2 per GPU, with no collective).
Prints seconds per edited image."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from PIL import Image
from bagel_b200 import synthetic
from bagel_b200.autoencoder import load_ae
from bagel_b200.config import SiglipVisionConfig
from bagel_b200.inferencer import InterleaveInferencer
from bagel_b200.modeling_utils import MLPconnector, PositionEmbedding
from bagel_b200.siglip_navit import SiglipVisionModel
from bagel_b200.transforms import ImageTransform
from oracle import fixtures   # random weights with the reference's key names (test infrastructure, not the product path)

dev = "cuda"
model = synthetic.build_random_bagel(device=dev, seed=0)
vcfg = SiglipVisionConfig(hidden_size=1152, intermediate_size=4304, num_hidden_layers=26, num_attention_heads=16,
                          num_channels=3, image_size=980, patch_size=14, rope=False)
vsd = fixtures.vit_state_dict(1152, 4304, 26, 16, 3584, max_side=70, seed=5)
vit = SiglipVisionModel(vcfg, dev)
vit.load_state_dict({k[len("vit_model."):]: v for k, v in vsd.items() if k.startswith("vit_model.")})
model.vit_model = vit; model.config.vit_config = vcfg; model.config.visual_und = True
model.vit_patch_size, model.vit_max_num_patch_per_side, model.vit_hidden_size = 14, 70, 1152
model.connector = MLPconnector(1152, 3584); model.connector.load(vsd, "connector.", dev)
model.vit_pos_embed = PositionEmbedding(70, 3584, dev)
vae, _ = load_ae(None, device=dev)
vae.load_state_dict(fixtures.vae_state_dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=16))
vae.sample = False
inf = InterleaveInferencer(model, vae, synthetic.RandomIdTokenizer(1), ImageTransform(1024, 512, 16), ImageTransform(980, 224, 14),
                           synthetic.NEW_TOKEN_IDS)
rs = np.random.RandomState(0)
img = Image.fromarray(rs.randint(0, 255, (1024, 1024, 3)).astype(np.uint8))
kw = dict(cfg_text_scale=4.0, cfg_img_scale=2.0, cfg_interval=[0.0, 1.0], timestep_shift=3.0, num_timesteps=50,
          cfg_renorm_min=0.0, cfg_renorm_type="text_channel")
for it in range(2):
    torch.manual_seed(it); torch.cuda.synchronize(); t0 = time.perf_counter()
    out = inf(image=img, text="64", **kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"[edit 1024^2, 3 CFG branches, 49 evals] run {it}: {dt:.2f} s per image ({out['image'].size})", flush=True)
torch.manual_seed(9); torch.cuda.synchronize(); t0 = time.perf_counter()
out = inf(image=img, text="64", enable_taylorseer=True, **kw)
torch.cuda.synchronize(); print(f"[edit, enable_taylorseer=True] {time.perf_counter() - t0:.2f} s per image", flush=True)
