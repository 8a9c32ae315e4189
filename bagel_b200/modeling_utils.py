"""Small heads around the LM: frozen 2-D sin-cos position tables, timestep embedder, ViT->LLM connector.
Reference: modeling/bagel/modeling_utils.py:24-144. Host-side table construction is numpy (load time only);
the per-step math runs through bagel_b200.ops kernels.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import ops

BF16 = torch.bfloat16


def sincos_2d_table(embed_dim: int, grid_size: int) -> torch.Tensor:
    """[grid_size**2, embed_dim] fp32. Row r*grid+c = [sincos(c) | sincos(r)], each half [sin(p w) | cos(p w)],
    w_k = 10000^(-k/(embed_dim/4)); float64 math then fp32 (reference :24-66 — note the column coordinate
    comes first because np.meshgrid(w, h) is called with w first)."""
    assert embed_dim % 4 == 0
    quarter = embed_dim // 4
    omega = 1.0 / (10000.0 ** (np.arange(quarter, dtype=np.float64) / quarter))
    coords = np.arange(grid_size, dtype=np.float32)
    col = np.tile(coords[None, :], (grid_size, 1)).reshape(-1)   # c varies fastest
    row = np.tile(coords[:, None], (1, grid_size)).reshape(-1)

    def enc(p):
        ang = np.einsum("m,d->md", p, omega)
        return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)

    table = np.concatenate([enc(col), enc(row)], axis=1)
    return torch.from_numpy(table).float()


class PositionEmbedding:
    """Frozen table looked up by flattened (row*max_side+col) ids (reference :127-144)."""

    def __init__(self, max_num_patch_per_side: int, hidden_size: int, device="cuda", dtype=BF16):
        """dtype: bf16 when the checkpoint is cast to bf16 (mode A), fp32 with fp32 master weights (mode B)."""
        self.max_num_patch_per_side = max_num_patch_per_side
        self.hidden_size = hidden_size
        self.dtype = dtype
        self.pos_embed = sincos_2d_table(hidden_size, max_num_patch_per_side).to(device, dtype).contiguous()

    def load(self, t: Optional[torch.Tensor]):
        if t is not None:
            self.pos_embed = t.to(self.pos_embed.device, self.dtype).contiguous()


def timestep_sinusoid(t: torch.Tensor, dim: int = 256, max_period: float = 10000.0) -> torch.Tensor:
    """[N] -> [N, dim] fp32, [cos | sin] halves, t used unscaled (reference :87-105). Plain torch on the
    caller's device: this is setup for a whole denoising run (all timesteps at once), not per-step work."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class TimestepEmbedder:
    """Linear(256,H) -> SiLU -> Linear(H,H) (reference :74-110); both GEMMs are bagel_gemm_bf16, SiLU fused."""

    def __init__(self, hidden_size: int, frequency_embedding_size: int = 256):
        self.hidden_size = hidden_size
        self.frequency_embedding_size = frequency_embedding_size
        self.w0 = self.b0 = self.w2 = self.b2 = None

    def load(self, sd: Dict[str, torch.Tensor], prefix: str, device):
        self.w0 = sd[prefix + "mlp.0.weight"].to(device, BF16).contiguous()
        self.b0 = sd[prefix + "mlp.0.bias"].to(device, BF16).contiguous()
        self.w2 = sd[prefix + "mlp.2.weight"].to(device, BF16).contiguous()
        self.b2 = sd[prefix + "mlp.2.bias"].to(device, BF16).contiguous()

    def __call__(self, t: torch.Tensor) -> torch.Tensor:
        """t [N] (fp32, device) -> [N, H] bf16."""
        freq = timestep_sinusoid(t.to(self.w0.device), self.frequency_embedding_size).to(BF16).contiguous()
        hid = ops.gemm(freq, self.w0, bias=self.b0, epilogue=ops.EPI_SILU)
        return ops.gemm(hid, self.w2, bias=self.b2)


class MLPconnector:
    """Linear -> GELU(tanh) -> Linear (reference :113-124)."""

    def __init__(self, in_dim: int, out_dim: int, hidden_act: str = "gelu_pytorch_tanh"):
        assert hidden_act == "gelu_pytorch_tanh"
        self.w1 = self.b1 = self.w2 = self.b2 = None

    def load(self, sd, prefix, device):
        self.w1 = sd[prefix + "fc1.weight"].to(device, BF16).contiguous()
        self.b1 = sd[prefix + "fc1.bias"].to(device, BF16).contiguous()
        self.w2 = sd[prefix + "fc2.weight"].to(device, BF16).contiguous()
        self.b2 = sd[prefix + "fc2.bias"].to(device, BF16).contiguous()

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        hid = ops.gemm(x.to(BF16).contiguous(), self.w1, bias=self.b1, epilogue=ops.EPI_GELU)
        return ops.gemm(hid, self.w2, bias=self.b2)
