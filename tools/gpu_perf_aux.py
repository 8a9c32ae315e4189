"""Timing of the non-LM rows of SURVEY.md §8 at BAGEL-7B shapes: FLUX VAE decode/encode at 1024^2 (vs cuDNN bf16 through
the same functional graph), SigLIP-so400m tower on 32 x 378^2 images (BASELINE configs[2] shape)."""
import sys, time, torch
sys.path.insert(0, ".")
from bagel_b200 import ops
from bagel_b200.autoencoder import AutoEncoder
from bagel_b200.config import AutoEncoderParams, SiglipVisionConfig
from bagel_b200.siglip_navit import SiglipVisionModel
from oracle import fixtures
from oracle import vae as ov

dev = "cuda"
# torch baseline = the oracle's functional graph on the GPU with CUDA-autocast semantics (group_norm runs in fp32)
import torch.nn.functional as F
ov.gn = lambda sd, name, x: F.group_norm(x.float(), 32, sd[name + ".weight"], sd[name + ".bias"], eps=1e-6)
def bench(fn, iters=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

# ---------------- VAE ----------------
sd = fixtures.vae_state_dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, seed=12)
ae = AutoEncoder(AutoEncoderParams(), dev); ae.load_state_dict(sd); ae.sample = False
sd_gpu = {k: v.to(dev) for k, v in sd.items()}
vc = ov.VaeConfig()
for side in (512, 1024):
    z = torch.randn(1, 16, side // 8, side // 8, device=dev)
    img = torch.rand(1, 3, side, side, device=dev) * 2 - 1
    t_dec = bench(lambda: ae.decode(z))
    t_enc = bench(lambda: ae.encode(img))
    print(f"[vae {side}^2 B=1] ours: decode {t_dec:.2f} ms, encode {t_enc:.2f} ms", flush=True)
    with torch.no_grad():
        t_dec_ref = bench(lambda: ov.decode(sd_gpu, vc, z))       # torch: cuDNN bf16 convs + ATen group_norm + SDPA
        t_enc_ref = bench(lambda: ov.encode(sd_gpu, vc, img))
        ref = ov.decode(sd_gpu, vc, z).float()
    got = ae.decode(z).float()
    fl_dec, fl_enc = 10.47e12 * (side / 1024) ** 2, 4.88e12 * (side / 1024) ** 2
    print(f"[vae {side}^2 B=1] decode ours {t_dec:.2f} ms ({fl_dec/t_dec/1e9:.0f} TFLOP/s) | torch/cuDNN {t_dec_ref:.2f} ms | "
          f"encode ours {t_enc:.2f} ms ({fl_enc/t_enc/1e9:.0f} TFLOP/s) | torch/cuDNN {t_enc_ref:.2f} ms | "
          f"max|ours-torch| {(got-ref).abs().max().item():.3e} (|x| max {ref.abs().max().item():.2f})", flush=True)

# ---------------- SigLIP so400m/14 (26 layers used) ----------------
vcfg = SiglipVisionConfig(hidden_size=1152, intermediate_size=4304, num_hidden_layers=26, num_attention_heads=16,
                          num_channels=3, image_size=980, patch_size=14, rope=False)
vsd = fixtures.vit_state_dict(1152, 4304, 26, 16, 3584, max_side=70, seed=5)
vit = SiglipVisionModel(vcfg, dev); vit.load_state_dict({k[len("vit_model."):]: v for k, v in vsd.items() if k.startswith("vit_model.")})
B, npatch = 32, 27 * 27
px = torch.randn(B * npatch, 588)
pos = (torch.arange(27)[:, None] * 70 + torch.arange(27)[None, :]).reshape(-1).repeat(B)
cu = torch.arange(0, (B + 1) * npatch, npatch, dtype=torch.int32)
t_vit = bench(lambda: vit(px.to(dev), pos, cu, npatch))
fl = B * npatch * 0.88e9 + 4.0 * B * npatch * npatch * 1152 * 26
print(f"[siglip so400m 26L] {B} x 378^2 ({B*npatch} tokens): {t_vit:.2f} ms = {B*npatch/t_vit*1e3:.0f} tokens/s, ~{fl/t_vit/1e9:.0f} TFLOP/s", flush=True)
