"""Host logic of the product (no GPU): the vectorised packers of bagel_b200.Bagel against the committed
reference outputs and against the oracle's loop restatement on random ragged inputs."""
import os
import random

import pytest
import torch
from safetensors.torch import load_file

import helpers
from oracle import bagel_flow as obf
from oracle import fixtures


@pytest.fixture(scope="module")
def model():
    return helpers.build_product_bagel(device="cpu", load=False)


def test_against_reference_fixture(model, golden_dir):
    g = load_file(os.path.join(golden_dir, "flow_tiny.safetensors"))
    gi, kv, rp = model.prepare_prompts([0, 0], [0, 0], helpers.PROMPTS, helpers.IntTokenizer(), helpers.NEW_TOKEN_IDS)
    for k, v in gi.items():
        assert torch.equal(v, g["prompts." + k]) and v.dtype == g["prompts." + k].dtype, k
    assert kv == g["prefill.kv_lens"].tolist() and rp == g["prefill.ropes"].tolist()
    torch.manual_seed(2)
    lat = model.prepare_vae_latent(kv, rp, helpers.IMAGE_SIZES, helpers.NEW_TOKEN_IDS)
    for k, v in lat.items():
        assert torch.equal(v, g["latent." + k]) and v.dtype == g["latent." + k].dtype, k
    c = model.prepare_vae_latent_cfg(kv, rp, helpers.IMAGE_SIZES)
    for k, v in c.items():
        assert torch.equal(v, g["cfg_img." + k]), k
    c0 = model.prepare_vae_latent_cfg([0, 0], [0, 0], helpers.IMAGE_SIZES)
    for k, v in c0.items():
        assert torch.equal(v, g["cfg_text." + k]), k


@pytest.mark.parametrize("seed", range(8))
def test_random_ragged_vs_oracle(model, seed):
    rnd = random.Random(seed)
    B = rnd.randint(1, 5)
    kv = [rnd.choice([0, 0, 1, 7, 66, 300]) for _ in range(B)]
    rp = [rnd.randint(0, 50) for _ in range(B)]
    prompts = [" ".join(str(rnd.randint(0, 999)) for _ in range(rnd.randint(0, 12))) for _ in range(B)]
    tok = helpers.IntTokenizer()
    fc = obf.FlowConfig(lm=fixtures.TINY_LM, max_latent_size=8)
    a, kv2, rp2 = model.prepare_prompts(kv, rp, prompts, tok, helpers.NEW_TOKEN_IDS)
    b, kv3, rp3 = obf.prepare_prompts(kv, rp, [tok.encode(p) for p in prompts], 1000, 1001)
    assert kv2 == kv3 and rp2 == rp3
    for k in b:
        assert torch.equal(a[k], b[k]) and a[k].dtype == b[k].dtype, k
    sizes = [(16 * rnd.randint(1, 8), 16 * rnd.randint(1, 8)) for _ in range(B)]
    torch.manual_seed(seed)
    a = model.prepare_vae_latent(kv2, rp2, sizes, helpers.NEW_TOKEN_IDS)
    torch.manual_seed(seed)
    b = obf.prepare_vae_latent(fc, kv2, rp2, sizes, 1002, 1003)
    for k in b:
        assert torch.equal(a[k], b[k]) and a[k].dtype == b[k].dtype, k
    a = model.prepare_vae_latent_cfg(kv2, rp2, sizes)
    b = obf.prepare_vae_latent_cfg(fc, kv2, rp2, sizes)
    for k in b:
        assert torch.equal(a[k], b[k]) and a[k].dtype == b[k].dtype, k


def test_start_tokens(model):
    out = model.prepare_start_tokens([3, 5], [3, 1], helpers.NEW_TOKEN_IDS)
    assert out["packed_start_tokens"].tolist() == [1000, 1000]
    assert out["packed_query_position_ids"].tolist() == [3, 1]
    assert out["key_values_lens"].tolist() == [3, 5] and out["key_values_lens"].dtype == torch.int32
    assert out["packed_key_value_indexes"].tolist() == [0, 1, 2, 3, 4, 5, 6, 7]  # reference quirk: no query slot


def test_position_ids_worked_example(model):
    # 32x32 and 32x48 images -> 2x2 and 2x3 latent tokens at max_latent_size 64
    from bagel_b200.bagel import get_flattened_position_ids_extrapolate as f
    assert f(32, 32, 16, 64).tolist() == [0, 1, 64, 65]
    assert f(32, 48, 16, 64).tolist() == [0, 1, 2, 64, 65, 66]


def test_vit_packer_vs_reference_fixture(golden_dir):
    m = helpers.build_product_bagel_with_vit(device="cpu", load=False)
    g = load_file(os.path.join(golden_dir, "vit_tiny.safetensors"))
    gi, kv, rp = m.prepare_vit_images([0, 0], [0, 0], fixtures.vit_images(), lambda im: im, helpers.NEW_TOKEN_IDS)
    for k, v in gi.items():
        assert torch.equal(v, g["vit_in." + k]) and v.dtype == g["vit_in." + k].dtype, k
    assert kv == g["vit_in.kv_lens"].tolist() and rp == g["vit_in.ropes"].tolist()


def test_vae_image_packer_vs_reference_fixture(golden_dir):
    from bagel_b200.config import AutoEncoderParams
    m = helpers.build_product_bagel(device="cpu", load=False, max_latent_size=16, vae_downsample=2)
    g = load_file(os.path.join(golden_dir, "vae_tiny.safetensors"))
    img, _ = fixtures.vae_inputs()
    gi, kv, rp = m.prepare_vae_images([0, 0], [0, 0], [img[0], img[1][:, :24, :32]], lambda im: im, helpers.NEW_TOKEN_IDS)
    for k, v in gi.items():
        if torch.is_tensor(v):
            assert torch.equal(v, g["vae_ctx." + k]) and v.dtype == g["vae_ctx." + k].dtype, k
    assert gi["patchified_vae_latent_shapes"] == [(8, 12), (6, 8)]
    assert kv == g["vae_ctx.kv_lens"].tolist() and rp == g["vae_ctx.ropes"].tolist()
