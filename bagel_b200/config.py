"""Configuration objects with the reference's field names (drop-in for the JSON files the reference loads:
llm_config.json -> Qwen2Config, vit_config.json -> SiglipVisionConfig; BagelConfig is built in code).

Reference: modeling/bagel/qwen2_navit.py:46-204 (Qwen2Config + qk_norm/layer_module/freeze_und),
modeling/bagel/siglip_navit.py:21-99 (SiglipVisionConfig + rope), modeling/bagel/bagel.py:27-54 (BagelConfig),
modeling/autoencoder.py:20-31 (AutoEncoderParams). No dependency on transformers.PretrainedConfig: these are
plain attribute bags that accept and keep unknown keys, so HF JSON files load unchanged.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional


class _Config:
    _defaults: Dict[str, Any] = {}

    def __init__(self, **kwargs):
        for k, v in self._defaults.items():
            setattr(self, k, v() if callable(v) else v)
        for k, v in kwargs.items():
            setattr(self, k, v)

    @classmethod
    def from_json_file(cls, path: str):
        with open(path, "r") as f:
            return cls(**json.load(f))

    @classmethod
    def from_dict(cls, d: Dict[str, Any]):
        return cls(**d)

    def to_dict(self) -> Dict[str, Any]:
        out = {}
        for k, v in self.__dict__.items():
            out[k] = v.to_dict() if isinstance(v, _Config) else v
        return out

    def __repr__(self):
        return f"{type(self).__name__}({self.to_dict()})"


class Qwen2Config(_Config):
    _defaults = dict(
        vocab_size=151936, hidden_size=4096, intermediate_size=22016, num_hidden_layers=32,
        num_attention_heads=32, num_key_value_heads=32, hidden_act="silu", max_position_embeddings=32768,
        initializer_range=0.02, rms_norm_eps=1e-6, use_cache=True, tie_word_embeddings=False,
        rope_theta=10000.0, rope_scaling=None, attention_dropout=0.0, pad_token_id=None,
        # BAGEL additions (qwen2_navit.py:196-204)
        qk_norm=True, layer_module="Qwen2DecoderLayer", freeze_und=False,
    )

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        if self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads
        if self.hidden_act != "silu":
            raise ValueError("bagel_b200 implements the SwiGLU (silu) MLP only")
        if self.rope_scaling not in (None, {}) and self.rope_scaling.get("rope_type", self.rope_scaling.get("type", "default")) != "default":
            raise ValueError("bagel_b200 implements default RoPE only (BAGEL ships rope_scaling=None)")

    @property
    def head_dim(self) -> int:
        return self.__dict__.get("_head_dim") or self.hidden_size // self.num_attention_heads

    @property
    def is_mot(self) -> bool:
        return "MoT" in self.layer_module


class SiglipVisionConfig(_Config):
    _defaults = dict(
        hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12, num_channels=3,
        image_size=224, patch_size=16, hidden_act="gelu_pytorch_tanh", layer_norm_eps=1e-6,
        attention_dropout=0.0, rope=True,
    )


@dataclass
class AutoEncoderParams:
    """FLUX VAE hyper-parameters (modeling/autoencoder.py:20-31; values fixed in load_ae :339-354)."""
    resolution: int = 256
    in_channels: int = 3
    downsample: int = 8
    ch: int = 128
    out_ch: int = 3
    ch_mult: List[int] = field(default_factory=lambda: [1, 2, 4, 4])
    num_res_blocks: int = 2
    z_channels: int = 16
    scale_factor: float = 0.3611
    shift_factor: float = 0.1159


class BagelConfig(_Config):
    _defaults = dict(
        visual_gen=True, visual_und=True, llm_config=None, vit_config=None, vae_config=None,
        latent_patch_size=2, max_latent_size=32, vit_max_num_patch_per_side=70,
        connector_act="gelu_pytorch_tanh", interpolate_pos=False, timestep_shift=1.0,
    )
