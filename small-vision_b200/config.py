"""Model and training configuration of the UMD auto-encoder hot path.

Restates the values the reference spreads over `_ViTAE`'s dataclass fields
(big_vision/models/ae.py:38-55), `decode_variant` (ae.py:200-218) and
`configs/ae_i1k.py:6-96`, as plain dataclasses (ml_collections is not needed on this path).
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Sequence, Tuple

_VARIANTS = {
    "S": dict(width=384, depth=12, dec_depth=4, num_heads=6),
    "B": dict(width=768, depth=12, dec_depth=4, num_heads=12),
    "L": dict(width=1024, depth=24, dec_depth=8, num_heads=16),
}


def decode_variant(variant: Optional[str]) -> dict:
  """ae.py:200-218 — "B" or "B/4" -> width/depth/dec_depth/num_heads (+ patch_size)."""
  if variant is None:
    return {}
  v, patch = variant, {}
  if "/" in variant:
    v, p = variant.split("/")
    patch = {"patch_size": (int(p), int(p))}
  if v not in _VARIANTS:
    raise KeyError(v)
  return {**_VARIANTS[v], **patch}


@dataclasses.dataclass(frozen=True)
class ModelConfig:
  """Fields and defaults of `_ViTAE` (ae.py:38-55)."""
  num_classes: Optional[int] = None
  channels: int = 3
  img_size: int = 64
  patch_size: Tuple[int, int] = (4, 4)
  width: int = 768
  depth: int = 12
  dec_depth: int = 4
  mlp_dim: Optional[int] = None
  num_heads: int = 12
  dropout: float = 0.0
  scan: bool = True
  remat_policy: str = "nothing_saveable"
  dtype_mm: str = "float32"
  adaln: bool = False
  cfg_dropout_rate: float = 0.1
  num_cls: int = 4
  no_decay_list: Sequence[str] = ("cls", "image_mask_embedding", "bias")
  # not a reference field: orientation of final_conv; True = flax default
  flip_final_conv: bool = True

  @property
  def patch(self) -> int:
    return int(self.patch_size[0])

  @property
  def grid(self) -> int:
    return self.img_size // self.patch

  @property
  def num_patches(self) -> int:
    return self.grid * self.grid

  @property
  def mlp(self) -> int:
    return self.mlp_dim or 4 * self.width

  def len_keep(self, mask_ratio: float) -> int:
    """ae.py:11 — int(L * (1 - mask_ratio)) in Python double arithmetic."""
    return int(self.num_patches * (1 - mask_ratio))


def make_model_config(*, variant=None, **kw) -> ModelConfig:
  """`Model(*, variant=None, **kw)` argument handling (ae.py:220-222)."""
  merged = {**decode_variant(variant), **kw}
  if "patch_size" in merged:
    merged["patch_size"] = tuple(int(x) for x in merged["patch_size"])
  if "no_decay_list" in merged:
    merged["no_decay_list"] = tuple(merged["no_decay_list"])
  if merged.get("dropout", 0.0) != 0.0:
    raise NotImplementedError("dropout > 0 is not used by any reference recipe (ae.py:48) and is not implemented")
  return ModelConfig(**merged)


@dataclasses.dataclass
class TrainConfig:
  """The training-step knobs of configs/ae_i1k.py:8-11,37-39,45-51,91-96 and train_ae.py:124-152."""
  batch_size: int = 1024
  no_noise_prob: float = 0.5
  mask_ratio: float = 0.375
  mask_ratio_no_noise: float = 0.75
  use_labels: bool = False
  beta_schedule: str = "cosine"      # 'linear' for latent diffusion (ae_i1k.py:45-50)
  timesteps: int = 1000
  diffusion_space: Tuple[int, int, int] = (64, 64, 3)
  peak_lr: float = 15e-5
  wd: float = 5e-2
  betas: Tuple[float, float] = (0.9, 0.95)
  clip_norm: float = 1.0
  total_epochs: int = 800
  ntrain_img: int = 1_268_355        # imagenet2012 train[:99%]
  mu_dtype: str = "bfloat16"
  ema_decay: Optional[float] = None  # 1e-4 * B/256 when use_labels (ae_i1k.py:31-33)
  total_steps: Optional[int] = None
  warmup_steps: Optional[int] = None

  def resolved(self) -> "TrainConfig":
    c = dataclasses.replace(self)
    steps_per_epoch = c.ntrain_img / c.batch_size
    if c.total_steps is None:
      c.total_steps = max(round(c.total_epochs * steps_per_epoch), 1)   # utils.steps rounds to nearest (utils.py:1059-1061)
    if c.warmup_steps is None:
      warmup_epochs = int(0.05 * c.total_epochs)           # ae_i1k.py:93
      c.warmup_steps = warmup_epochs * c.ntrain_img // c.batch_size   # train_ae.py:137
    return c

  @property
  def scaled_peak_lr(self) -> float:
    return self.peak_lr * self.batch_size / 256            # train_ae.py:136


def warmup_cosine_lr(count: int, *, peak: float, warmup_steps: int, decay_steps: int, init_value: float = 0.0,
                     end_value: float = 0.0) -> float:
  """optax.warmup_cosine_decay_schedule evaluated on the host at the pre-increment step count."""
  import math
  if warmup_steps > 0 and count < warmup_steps:
    return init_value + (peak - init_value) * (max(count, 0) / warmup_steps)
  span = max(decay_steps - warmup_steps, 1)
  c = min(max(count - warmup_steps, 0), span)
  cos = 0.5 * (1.0 + math.cos(math.pi * c / span))
  alpha = (end_value / peak) if peak else 0.0
  return peak * ((1.0 - alpha) * cos + alpha)
