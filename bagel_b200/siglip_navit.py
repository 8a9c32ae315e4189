"""SigLIP NaViT vision encoder — host side (reference: modeling/bagel/siglip_navit.py:145-402).

Same call surface as the reference's SiglipVisionModel: `model(packed_pixel_values, packed_flattened_position_ids,
cu_seqlens, max_seqlen)` -> [tokens, hidden]; `.vision_model.embeddings.convert_conv2d_to_linear(cfg)` is accepted
(weights are always held in the linear layout). State-dict keys follow the reference.

Kernel mapping per encoder layer: LayerNorm -> fused QKV GEMM (+bias) -> packed varlen attention -> out-proj GEMM
with residual epilogue -> LayerNorm -> fc1 GEMM with GELU(tanh) epilogue -> fc2 GEMM with residual epilogue.
The tower's head_dim is 72 (1152/16), which is neither a TMA/UMMA-friendly width nor a multiple of 16: the fused
QKV weight is laid out with every head padded to 128 rows (zero weights/bias for the padding), so the attention
kernel runs its d=128 path (softmax scale 72^-0.5 passed explicitly) and the out-projection weight carries zero
columns for the padding. head_dim 64/128 towers run unpadded.

`rope=True` (2-D RoPE, siglip_navit.py:102-142, 224-230; disabled in every shipped inference config, app.py:45,
eval/vlm/utils.py:37): the tower then has no learned position table; q/k heads are rotated in place in the fused QKV
buffer (bagel_siglip_rope2d_bf16: row table on the first half of a head, column table on the second half, fp32 tables
built at construction exactly like RotaryEmbedding2D) between the QKV GEMM and the attention kernel.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops
from .config import SiglipVisionConfig

BF16 = torch.bfloat16


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _rope2d_tables(dim: int, max_h: int, max_w: int, base: float = 10000.0):
    """cos_h, sin_h, cos_w, sin_w fp32 [max_h*max_w, dim] — the buffers of the reference's RotaryEmbedding2D (:102-127),
    same expressions on the host (construction time only)."""
    inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))
    grid_h = torch.arange(0, max_h).to(inv_freq.dtype)[:, None].repeat(1, max_w)
    grid_w = torch.arange(0, max_w).to(inv_freq.dtype)[None, :].repeat(max_h, 1)
    out = []
    for grid in (grid_h, grid_w):
        freqs = grid[..., None] * inv_freq[None, None, :]
        emb = torch.cat((freqs, freqs), dim=-1).flatten(0, 1)
        out += [emb.cos(), emb.sin()]
    return tuple(out)


class _Embeddings:
    def __init__(self, owner):
        self._owner = owner

    def convert_conv2d_to_linear(self, config=None, meta=False):
        """No-op: the patch embedding is always stored as Linear(C*p*p -> hidden) in (p, q, c) column order
        (reference :167-180 performs `W.permute(0,2,3,1).reshape(...)`; load_state_dict does it for 4-D weights)."""
        return None


class SiglipVisionTransformer:
    def __init__(self, config: SiglipVisionConfig, device="cuda"):
        self.config = config
        self.device = torch.device(device)
        self.rope = bool(getattr(config, "rope", False))
        self.embeddings = _Embeddings(self)
        H, nh = config.hidden_size, config.num_attention_heads
        self.head_dim = H // nh
        self.head_pad = self.head_dim if self.head_dim in (64, 128) else (64 if self.head_dim < 64 else 128)
        if self.head_dim > 128:
            raise NotImplementedError("SigLIP head_dim > 128")
        self.patch_dim = config.num_channels * config.patch_size ** 2
        self.patch_dim_pad = _pad8(self.patch_dim)
        self.layers = []
        self.w = {}
        self._ws: Dict[str, torch.Tensor] = {}
        if self.rope:
            if self.head_dim % 4:
                raise ValueError("SigLIP 2-D RoPE needs head_dim % 4 == 0")
            side = config.image_size // config.patch_size
            self.rope_tables = tuple(t.to(self.device).contiguous() for t in _rope2d_tables(self.head_dim // 2, side, side))

    def _buf(self, name, rows, cols):
        t = self._ws.get(name)
        if t is None or t.shape[0] < rows or t.shape[1] != cols:
            t = torch.empty((rows, cols), dtype=BF16, device=self.device)
            self._ws[name] = t
        return t[:rows]

    # ---- weights ------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = ""):
        cfg, dev = self.config, self.device
        H, nh, d, dp = cfg.hidden_size, cfg.num_attention_heads, self.head_dim, self.head_pad

        def get(k):
            return sd[prefix + k].to(dev, BF16)

        pw = get("embeddings.patch_embedding.weight")
        if pw.dim() == 4:  # conv layout [H, C, p, p] -> linear [(p, q, c)]
            pw = pw.permute(0, 2, 3, 1).reshape(H, self.patch_dim)
        wpad = torch.zeros((H, self.patch_dim_pad), dtype=BF16, device=dev)
        wpad[:, : self.patch_dim] = pw
        self.w["patch_w"] = wpad
        self.w["patch_b"] = get("embeddings.patch_embedding.bias").contiguous()
        if not self.rope:   # rope=True towers have no learned position table (:164-165)
            self.w["pos"] = get("embeddings.position_embedding.weight").contiguous()
        self.layers = []
        for li in range(cfg.num_hidden_layers):
            p = f"encoder.layers.{li}."
            L = {}
            wq, wk, wv = (get(p + f"self_attn.{n}_proj.weight") for n in "qkv")
            bq, bk, bv = (get(p + f"self_attn.{n}_proj.bias") for n in "qkv")
            wqkv = torch.zeros((3, nh, dp, H), dtype=BF16, device=dev)
            bqkv = torch.zeros((3, nh, dp), dtype=BF16, device=dev)
            for i, (w_, b_) in enumerate(((wq, bq), (wk, bk), (wv, bv))):
                wqkv[i, :, :d] = w_.reshape(nh, d, H)
                bqkv[i, :, :d] = b_.reshape(nh, d)
            L["wqkv"] = wqkv.reshape(3 * nh * dp, H).contiguous()
            L["bqkv"] = bqkv.reshape(-1).contiguous()
            wo = torch.zeros((H, nh, dp), dtype=BF16, device=dev)
            wo[:, :, :d] = get(p + "self_attn.out_proj.weight").reshape(H, nh, d)
            L["wo"] = wo.reshape(H, nh * dp).contiguous()
            L["bo"] = get(p + "self_attn.out_proj.bias").contiguous()
            for n in ("layer_norm1", "layer_norm2"):
                L[n + "_w"] = get(p + n + ".weight").contiguous()
                L[n + "_b"] = get(p + n + ".bias").contiguous()
            L["fc1_w"] = get(p + "mlp.fc1.weight").contiguous()
            L["fc1_b"] = get(p + "mlp.fc1.bias").contiguous()
            L["fc2_w"] = get(p + "mlp.fc2.weight").contiguous()
            L["fc2_b"] = get(p + "mlp.fc2.bias").contiguous()
            self.layers.append(L)
        self.w["post_w"] = get("post_layernorm.weight").contiguous()
        self.w["post_b"] = get("post_layernorm.bias").contiguous()

    # ---- forward ------------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, packed_pixel_values, packed_flattened_position_ids, cu_seqlens, max_seqlen):
        cfg, dev = self.config, self.device
        H, nh, dp = cfg.hidden_size, cfg.num_attention_heads, self.head_pad
        n = packed_pixel_values.shape[0]
        eps = cfg.layer_norm_eps
        px = torch.zeros((n, self.patch_dim_pad), dtype=BF16, device=dev)
        px[:, : self.patch_dim] = packed_pixel_values.to(dev, BF16)
        pos = torch.as_tensor(packed_flattened_position_ids).to(dev, torch.int64).contiguous()
        cu = torch.as_tensor(cu_seqlens).to(dev, torch.int32).contiguous()
        xa, xb = self._buf("xa", n, H), self._buf("xb", n, H)
        h = self._buf("h", n, H)
        qkv = self._buf("qkv", n, 3 * nh * dp)
        att = self._buf("att", n, nh * dp)
        mid = self._buf("mid", n, cfg.intermediate_size)
        # patch embed + learned position embedding (siglip_navit.py:190-193)
        if self.rope:
            ops.gemm(px, self.w["patch_w"], bias=self.w["patch_b"], out=xa)
        else:
            ops.gemm(px, self.w["patch_w"], bias=self.w["patch_b"], out=h)
            ops.latent_embed_add(h, None, self.w["pos"], pos, xa, None)
        scale = float(self.head_dim) ** -0.5
        q3 = qkv.view(n, 3 * nh, dp)
        for L in self.layers:
            ops.layernorm(xa, L["layer_norm1_w"], L["layer_norm1_b"], eps, out=h)
            ops.gemm(h, L["wqkv"], bias=L["bqkv"], out=qkv)
            if self.rope:   # q heads then k heads are the first 2*nh heads of the fused buffer
                ops.siglip_rope2d(qkv, 2 * nh, dp, self.head_dim, pos, *self.rope_tables)
            ops.attn_varlen(q3[:, :nh], q3[:, nh:2 * nh], q3[:, 2 * nh:], cu, cu, int(max_seqlen), int(max_seqlen), False,
                            softmax_scale=scale, out=att.view(n, nh, dp))
            ops.gemm(att, L["wo"], bias=L["bo"], resid=xa, epilogue=ops.EPI_RESID, out=xb)
            ops.layernorm(xb, L["layer_norm2_w"], L["layer_norm2_b"], eps, out=h)
            ops.gemm(h, L["fc1_w"], bias=L["fc1_b"], epilogue=ops.EPI_GELU, out=mid)
            ops.gemm(mid, L["fc2_w"], bias=L["fc2_b"], resid=xb, epilogue=ops.EPI_RESID, out=xa)
        out = torch.empty((n, H), dtype=BF16, device=dev)
        ops.layernorm(xa, self.w["post_w"], self.w["post_b"], eps, out=out)
        return out


class SiglipVisionModel:
    """Reference SiglipVisionModel (siglip_navit.py:374-402): `.vision_model`, callable with the packed inputs."""

    def __init__(self, config: SiglipVisionConfig, device="cuda"):
        self.config = config
        self.vision_model = SiglipVisionTransformer(config, device)

    def eval(self):
        return self

    def load_state_dict(self, sd, strict=True):
        self.vision_model.load_state_dict(sd, prefix="vision_model.")
        return self

    def __call__(self, packed_pixel_values, packed_flattened_position_ids, cu_seqlens, max_seqlen):
        return self.vision_model(packed_pixel_values, packed_flattened_position_ids, cu_seqlens, max_seqlen)

    forward = __call__
