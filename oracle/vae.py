"""TEST INFRASTRUCTURE — CPU restatement of the FLUX VAE (reference: modeling/autoencoder.py:34-325), functional
over a state dict with the reference's key names. Semantics are those of the reference run with fp32 parameters
under `torch.autocast("cpu", bfloat16)` (eval-driver style): conv2d and scaled_dot_product_attention run in bf16,
group_norm / swish / adds run in the dtype of their inputs. Do not call under autocast (casts are explicit)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import torch
import torch.nn.functional as F

from . import qwen2_mot as om


@dataclass
class VaeConfig:
    ch: int = 128
    ch_mult: List[int] = field(default_factory=lambda: [1, 2, 4, 4])
    num_res_blocks: int = 2
    z_channels: int = 16
    in_channels: int = 3
    out_ch: int = 3
    scale_factor: float = 0.3611
    shift_factor: float = 0.1159


def _ac():
    return om._AUTOCAST[0]


def conv(sd, name, x, stride=1, padding=0):
    dt = _ac()
    return F.conv2d(x.to(dt), sd[name + ".weight"].to(dt), sd[name + ".bias"].to(dt), stride=stride, padding=padding)


def gn(sd, name, x):
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], eps=1e-6)


def swish(x):
    return x * torch.sigmoid(x)


def resnet_block(sd, name, x):
    """autoencoder.py:68-95"""
    h = conv(sd, name + ".conv1", swish(gn(sd, name + ".norm1", x)), padding=1)
    h = conv(sd, name + ".conv2", swish(gn(sd, name + ".norm2", h)), padding=1)
    if (name + ".nin_shortcut.weight") in sd:
        x = conv(sd, name + ".nin_shortcut", x)
    return x + h


def attn_block(sd, name, x):
    """autoencoder.py:38-65"""
    h = gn(sd, name + ".norm", x)
    q, k, v = conv(sd, name + ".q", h), conv(sd, name + ".k", h), conv(sd, name + ".v", h)
    b, c, hh, ww = q.shape
    seq = lambda t: t.reshape(b, c, hh * ww).transpose(1, 2).reshape(b, 1, hh * ww, c).contiguous()
    dt = _ac()
    o = F.scaled_dot_product_attention(seq(q).to(dt), seq(k).to(dt), seq(v).to(dt))
    o = o.reshape(b, hh * ww, c).transpose(1, 2).reshape(b, c, hh, ww)
    return x + conv(sd, name + ".proj_out", o)


def encoder(sd, vc: VaeConfig, x, pfx="encoder"):
    """autoencoder.py:122-193"""
    h = conv(sd, pfx + ".conv_in", x, padding=1)
    n = len(vc.ch_mult)
    for lvl in range(n):
        for i in range(vc.num_res_blocks):
            h = resnet_block(sd, f"{pfx}.down.{lvl}.block.{i}", h)
        if lvl != n - 1:
            h = conv(sd, f"{pfx}.down.{lvl}.downsample.conv", F.pad(h, (0, 1, 0, 1), mode="constant", value=0), stride=2)
    h = resnet_block(sd, pfx + ".mid.block_1", h)
    h = attn_block(sd, pfx + ".mid.attn_1", h)
    h = resnet_block(sd, pfx + ".mid.block_2", h)
    return conv(sd, pfx + ".conv_out", swish(gn(sd, pfx + ".norm_out", h)), padding=1)


def decoder(sd, vc: VaeConfig, z, pfx="decoder"):
    """autoencoder.py:196-272"""
    h = conv(sd, pfx + ".conv_in", z, padding=1)
    h = resnet_block(sd, pfx + ".mid.block_1", h)
    h = attn_block(sd, pfx + ".mid.attn_1", h)
    h = resnet_block(sd, pfx + ".mid.block_2", h)
    for lvl in reversed(range(len(vc.ch_mult))):
        for i in range(vc.num_res_blocks + 1):
            h = resnet_block(sd, f"{pfx}.up.{lvl}.block.{i}", h)
        if lvl != 0:
            h = conv(sd, f"{pfx}.up.{lvl}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"), padding=1)
    return conv(sd, pfx + ".conv_out", swish(gn(sd, pfx + ".norm_out", h)), padding=1)


def encode(sd, vc: VaeConfig, x, noise=None):
    """autoencoder.py:275-287, 315-318 — DiagonalGaussian draws mean + std * eps (eps given explicitly here)."""
    mean, logvar = torch.chunk(encoder(sd, vc, x), 2, dim=1)
    z = mean if noise is None else mean + torch.exp(0.5 * logvar) * noise
    return vc.scale_factor * (z - vc.shift_factor)


def decode(sd, vc: VaeConfig, z):
    """autoencoder.py:320-322"""
    return decoder(sd, vc, z / vc.scale_factor + vc.shift_factor)
