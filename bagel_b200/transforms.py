"""Host-side image transforms at the boundary (reference: data/transforms.py:15-115, data/data_utils.py:118-127).
PIL in, normalised CHW float tensor out; bicubic antialiased resize to a stride multiple within
[min_size, max_size] and a pixel budget. Pure CPU preprocessing — not on the GPU hot path."""
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
