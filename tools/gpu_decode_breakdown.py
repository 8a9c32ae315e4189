"""Per-kernel breakdown of one text-decode step (run under `ncu --profile-from-start off --metrics gpu__time_duration.sum`)."""
import sys, torch
sys.path.insert(0, ".")
from bagel_b200 import synthetic
from bagel_b200.qwen2_navit import NaiveCache
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
model = synthetic.build_random_bagel(device="cuda", seed=0, num_layers=layers)
model.use_cuda_graph = False
B = 32
tok = synthetic.RandomIdTokenizer(1)
gi, kv, rp = model.prepare_prompts([0] * B, [0] * B, ["1243"] * B, tok, synthetic.NEW_TOKEN_IDS)
cache = model.forward_cache_update_text(NaiveCache(layers), **gi)
gs = model.prepare_start_tokens(kv, rp, synthetic.NEW_TOKEN_IDS)
torch.cuda.synchronize()
print("DECODE_BEGIN", flush=True)
torch.cuda.profiler.start()
toks = model.generate_text(past_key_values=cache, max_length=3, do_sample=False, **gs)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("DECODE_END", toks.shape, flush=True)
