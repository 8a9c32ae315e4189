"""Image transforms at the boundary (reference: data/transforms.py:15-115, data/data_utils.py:118-127).
PIL in, normalised CHW float tensor out; bicubic antialiased resize to a stride multiple within
[min_size, max_size] and a pixel budget.

`ImageTransform` is the host path (PIL + torch on the CPU, exactly the reference's arithmetic). `DeviceImageTransform`
keeps the same interface but uploads the uint8 pixels once and does the resize (bit-exact re-implementation of Pillow's
8-bit bicubic resampler), ToTensor + Normalize and — through `patches()` — the ViT patchify on the GPU
(bagel_b200/csrc/image.cu): its outputs are CUDA tensors bit-identical to the host path's."""
from __future__ import annotations

import numpy as np
import torch


def pil_img2rgb(image):
    """Composite transparency over white, then RGB (reference data/data_utils.py:118-127)."""
    from PIL import Image
    if image.mode == "RGBA" or image.info.get("transparency", None) is not None:
        image = image.convert("RGBA")
        canvas = Image.new(mode="RGB", size=image.size, color=(255, 255, 255))
        canvas.paste(image, mask=image.split()[3])
        return canvas
    return image.convert("RGB")


class MaxLongEdgeMinShortEdgeResize:
    def __init__(self, max_size: int, min_size: int, stride: int, max_pixels: int):
        self.max_size, self.min_size, self.stride, self.max_pixels = max_size, min_size, stride, max_pixels

    def _snap(self, v: float) -> int:
        return max(self.stride, int(round(v / self.stride) * self.stride))

    def _scaled(self, w, h, scale):
        return self._snap(round(w * scale)), self._snap(round(h * scale))

    def target_size(self, width: int, height: int, img_num: int = 1):
        scale = min(self.max_size / max(width, height), 1.0)
        scale = max(scale, self.min_size / min(width, height))
        w, h = self._scaled(width, height, scale)
        if w * h > self.max_pixels / img_num:
            w, h = self._scaled(w, h, self.max_pixels / img_num / (w * h))
        if max(w, h) > self.max_size:
            w, h = self._scaled(w, h, self.max_size / max(w, h))
        return w, h

    def __call__(self, img, img_num: int = 1):
        from PIL import Image
        w, h = self.target_size(*img.size, img_num=img_num)
        return img.resize((w, h), resample=Image.BICUBIC, reducing_gap=None)


class ImageTransform:
    def __init__(self, max_image_size, min_image_size, image_stride, max_pixels=14 * 14 * 9 * 1024,
                 image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5)):
        self.stride = image_stride
        self.resize_transform = MaxLongEdgeMinShortEdgeResize(max_image_size, min_image_size, image_stride, max_pixels)
        self.mean = torch.tensor(image_mean).view(3, 1, 1)
        self.std = torch.tensor(image_std).view(3, 1, 1)

    def __call__(self, img, img_num: int = 1) -> torch.Tensor:
        img = self.resize_transform(img, img_num=img_num)
        t = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).float() / 255.0
        return (t - self.mean) / self.std


# ---------------------------------------------------------------------------------------------------------------------
# device-side path
# ---------------------------------------------------------------------------------------------------------------------
def pil_bicubic_coeffs(in_size: int, out_size: int):
    """Fixed-point taps and windows of Pillow's 8-bit bicubic resampler for one axis (src/libImaging/Resample.c:
    precompute_coeffs + normalize_coeffs_8bpc, filter a = -0.5, support 2 stretched by max(scale, 1) for antialiasing).
    Returns (kk int32 [out_size, ksize], bounds int32 [out_size, 2] = (xmin, n)). Pure double-precision host arithmetic
    in the same operation order as Pillow, so the integers are identical."""
    import math
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 2.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = (xx + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)          # C (int) cast: truncation (values >= -2)
    xmin = np.where(center - support + 0.5 < 0, 0, xmin)
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size)
    n = (xmax - xmin).astype(np.int64)
    k = np.arange(ksize, dtype=np.float64)[None, :]
    x = np.abs((k + xmin[:, None] - center[:, None] + 0.5) * (1.0 / fscale))
    a = -0.5
    w = np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1, np.where(x < 2.0, (((x - 5) * x + 8) * x - 4) * a, 0.0))
    w = np.where(k < n[:, None], w, 0.0)
    ww = np.zeros(out_size, dtype=np.float64)
    for j in range(ksize):            # Pillow accumulates the taps left to right
        ww = ww + w[:, j]
    w = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    kk = np.where(w < 0, np.trunc(-0.5 + w * (1 << 22)), np.trunc(0.5 + w * (1 << 22))).astype(np.int32)
    bounds = np.stack([xmin, n], axis=1).astype(np.int32)
    return np.ascontiguousarray(kk), np.ascontiguousarray(bounds)


class DeviceImageTransform(ImageTransform):
    """ImageTransform whose pixel work runs on the GPU. `__call__` -> CUDA fp32 [3, H, W]; `patches(img, p)` -> CUDA fp32
    [(H/p)*(W/p), p*p*3] (the `packed_vit_tokens` rows). Bit-identical to the host path (tests/test_gpu_image.py)."""

    def __init__(self, max_image_size, min_image_size, image_stride, max_pixels=14 * 14 * 9 * 1024,
                 image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5), device="cuda"):
        super().__init__(max_image_size, min_image_size, image_stride, max_pixels, image_mean, image_std)
        self.device = torch.device(device)
        self._mean = [float(v) for v in image_mean]
        self._std = [float(v) for v in image_std]
        self._coeffs = {}

    def _taps(self, n_in: int, n_out: int):
        key = (n_in, n_out)
        if key not in self._coeffs:
            kk, b = pil_bicubic_coeffs(n_in, n_out)
            self._coeffs[key] = (torch.from_numpy(kk).to(self.device), torch.from_numpy(b).to(self.device), kk.shape[1])
        return self._coeffs[key]

    def resized_u8(self, img, img_num: int = 1) -> torch.Tensor:
        """PIL image -> uint8 [Ho, Wo, 3] on the device, resized exactly as `resize_transform` (Pillow) would."""
        from . import ops
        if img.mode != "RGB":
            img = img.convert("RGB")
        wi, hi = img.size
        wo, ho = self.resize_transform.target_size(wi, hi, img_num=img_num)
        src = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).pin_memory().to(self.device, non_blocking=True)
        if (wo, ho) == (wi, hi):
            return src
        return ops.image_resize_bicubic_u8(src, ho, wo, self._taps(wi, wo) if wo != wi else None,
                                           self._taps(hi, ho) if ho != hi else None)

    def __call__(self, img, img_num: int = 1) -> torch.Tensor:
        from . import ops
        return ops.image_normalize_u8(self.resized_u8(img, img_num), self._mean, self._std, patch=0)

    def patches(self, img, patch_size: int, img_num: int = 1) -> torch.Tensor:
        from . import ops
        return ops.image_normalize_u8(self.resized_u8(img, img_num), self._mean, self._std, patch=patch_size)
