import sys, os, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import helpers
from oracle import fixtures
from bagel_b200.inferencer import InterleaveInferencer
from bagel_b200.transforms import ImageTransform
from safetensors.torch import load_file
gold = load_file("tests/golden/inferencer_tiny.safetensors")
TEXT = open("tests/test_gpu_inferencer.py").read().split('TEXT = ')[1].split('\n')[0]
TEXT = eval(TEXT)
KW = dict(num_timesteps=4, timestep_shift=3.0, cfg_text_scale=4.0, cfg_img_scale=1.5, cfg_interval=[0.4, 1.0], cfg_renorm_min=0.0, cfg_renorm_type="global")
def mk():
    model = helpers.build_product_bagel_with_vit(fixtures.TINY_LM, "cuda", max_latent_size=16, vae_downsample=2)
    vae = helpers.tiny_vae("cuda"); vae.sample = False
    return InterleaveInferencer(model, vae, fixtures.ToyTokenizer(), ImageTransform(64, 32, 4), ImageTransform(112, 56, 14), helpers.NEW_TOKEN_IDS)
ref = bytes(gold["think.text"].tolist()).decode()
inf = mk()
for trial in range(3):
    torch.manual_seed(24)
    r = inf(text=TEXT, think=True, max_think_token_n=5, do_sample=False, image_shapes=(32, 48), **KW)
    print("fresh model trial", trial, "think:", r["text"], "| ref:", ref, flush=True)
inf = mk()
torch.manual_seed(21); inf(text=TEXT, image_shapes=(32, 48), **KW)
torch.manual_seed(24)
r = inf(text=TEXT, think=True, max_think_token_n=5, do_sample=False, image_shapes=(32, 48), **KW)
print("after t2i: think:", r["text"], flush=True)
torch.manual_seed(23)
r = inf(image=fixtures.inferencer_image(), text=TEXT, understanding_output=True, max_think_token_n=6, do_sample=False)
print("und:", r["text"], "| ref:", bytes(gold["und.text"].tolist()).decode(), flush=True)
torch.manual_seed(24)
r = inf(text=TEXT, think=True, max_think_token_n=5, do_sample=False, image_shapes=(32, 48), **KW)
print("after und: think:", r["text"], flush=True)
inf.model.use_cuda_graph = False
torch.manual_seed(24)
r = inf(text=TEXT, think=True, max_think_token_n=5, do_sample=False, image_shapes=(32, 48), **KW)
print("no graph: think:", r["text"], flush=True)
