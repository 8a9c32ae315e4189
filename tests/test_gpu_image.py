"""-m gpu: SURVEY.md §8(f) #3 — the device-side image path (bagel_b200/csrc/image.cu, transforms.DeviceImageTransform).
Integer / byte work: BIT-EXACT against Pillow's own bicubic resize (the reference's data/transforms.py:15-115 calls
PIL.Image.resize) and against the host ImageTransform + patchify (data/data_utils.py:43-50)."""
import numpy as np
import pytest
import torch
from PIL import Image

from bagel_b200 import ops
from bagel_b200.bagel import patchify
from bagel_b200.transforms import DeviceImageTransform, ImageTransform, pil_bicubic_coeffs

pytestmark = pytest.mark.gpu

SIZES = [  # (h, w) -> (ho, wo): up, down (antialiased: wide windows), one axis only, identity, extreme ratios
    ((40, 56), (56, 80)), ((378, 378), (392, 378)), ((100, 67), (42, 28)), ((64, 64), (64, 64)), ((301, 200), (980, 644)),
    ((1024, 768), (512, 384)), ((33, 47), (112, 154)), ((2000, 1500), (224, 168)), ((17, 900), (1024, 64)),
]


@pytest.mark.parametrize("src,dst", SIZES)
def test_resize_bit_exact_vs_pillow(src, dst):
    (h, w), (ho, wo) = src, dst
    a = np.random.RandomState(h * 7 + w).randint(0, 256, (h, w, 3)).astype(np.uint8)
    ref = np.asarray(Image.fromarray(a).resize((wo, ho), resample=Image.BICUBIC, reducing_gap=None))

    def taps(n_in, n_out):
        if n_in == n_out:
            return None
        kk, b = pil_bicubic_coeffs(n_in, n_out)
        return torch.from_numpy(kk).cuda(), torch.from_numpy(b).cuda(), kk.shape[1]

    got = ops.image_resize_bicubic_u8(torch.from_numpy(a).cuda(), ho, wo, taps(w, wo), taps(h, ho))
    assert np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.parametrize("size,args", [((40, 56), (64, 32, 4)), ((301, 200), (980, 224, 14)), ((378, 378), (980, 378, 14)),
                                       ((700, 1100), (1024, 512, 16)), ((90, 64), (112, 56, 14))])
def test_device_transform_equals_host_transform(size, args):
    h, w = size
    img = Image.fromarray(np.random.RandomState(h + w).randint(0, 256, (h, w, 3)).astype(np.uint8))
    host, dev = ImageTransform(*args), DeviceImageTransform(*args)
    a, b = host(img), dev(img)
    assert b.is_cuda and b.dtype == torch.float32 and tuple(a.shape) == tuple(b.shape)
    assert torch.equal(a, b.cpu())
    p = args[2]
    assert torch.equal(patchify(a, p), dev.patches(img, p).cpu())
    # non-default statistics exercise the three fp32 roundings of ToTensor + Normalize
    kw = dict(image_mean=(0.485, 0.456, 0.406), image_std=(0.229, 0.224, 0.225))
    assert torch.equal(ImageTransform(*args, **kw)(img), DeviceImageTransform(*args, **kw)(img).cpu())


def test_prepare_vit_and_vae_images_with_device_transform():
    """The packers accept the device transform: same dict (values on the GPU) as with the host transform."""
    import helpers
    from oracle import fixtures
    model = helpers.build_product_bagel_with_vit(fixtures.TINY_LM, "cuda", max_latent_size=16, vae_downsample=2)
    imgs = [fixtures.inferencer_image(3, 40, 56), fixtures.inferencer_image(4, 90, 64)]
    args = (112, 56, 14)
    gi_h, kv_h, rp_h = model.prepare_vit_images([0, 0], [0, 0], imgs, ImageTransform(*args), helpers.NEW_TOKEN_IDS)
    gi_d, kv_d, rp_d = model.prepare_vit_images([0, 0], [0, 0], imgs, DeviceImageTransform(*args), helpers.NEW_TOKEN_IDS)
    assert kv_h == kv_d and rp_h == rp_d and gi_d["packed_vit_tokens"].is_cuda
    for k in gi_h:
        assert torch.equal(gi_h[k], gi_d[k].cpu()), k
    args = (64, 32, 4)
    gv_h, _, _ = model.prepare_vae_images([0, 0], [0, 0], imgs, ImageTransform(*args), helpers.NEW_TOKEN_IDS)
    gv_d, _, _ = model.prepare_vae_images([0, 0], [0, 0], imgs, DeviceImageTransform(*args), helpers.NEW_TOKEN_IDS)
    assert gv_d["padded_images"].is_cuda and torch.equal(gv_h["padded_images"], gv_d["padded_images"].cpu())
    assert gv_h["patchified_vae_latent_shapes"] == gv_d["patchified_vae_latent_shapes"]
