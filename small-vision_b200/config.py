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
  # not a reference field: dtype of the residual stream between the blocks during training.  "float32" (default) keeps
  # the stream of the reference's default dtype_mm="float32"; "bfloat16" is the stream of its dtype_mm="bfloat16" flow
  # (ae.py:51,100: Dense / Conv outputs and x + y are bf16, LayerNorm statistics fp32)
  residual_dtype: str = "float32"
  # ... and of the gradient of that stream in the backward pass ("bfloat16" needs residual_dtype="bfloat16": JAX cotangents
  # take the dtype of their primals, so this is the backward of the same dtype_mm="bfloat16" flow)
  grad_stream_dtype: str = "float32"

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
  if merged.get("grad_stream_dtype", "float32") == "bfloat16" and merged.get("residual_dtype", "float32") != "bfloat16":
    raise ValueError("grad_stream_dtype='bfloat16' requires residual_dtype='bfloat16'")
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


# ------------------------------------------------------------------------------------------------------------------
# The reference's launcher-facing recipe: `--config big_vision/configs/ae_i1k.py:variant=B/4,batch_size=4096,...`
# ------------------------------------------------------------------------------------------------------------------
_ARG_SPEC = dict(variant="B/4", scan=True, fsdp=False, batch_size=1024, use_labels=False, mask_ratio=0.375, no_noise_prob=0.5,
                 mask_ratio_no_noise=0.75, finetune=False, lr=15e-5, wd=5e-2, beta2=0.95, size=64, adaln=True, epochs=800,
                 area_min=80, use_preprocessed_latents=False, latent_diffusion=False, wandb_mode="online", save_ckpt=True)


def parse_arg(arg: Optional[str]) -> dict:
  """The option string of configs/ae_i1k.py:8-11 (`name=value,flag,...`; configs/common.py:29-103): values are converted
  with the type of their default, booleans strictly from 'true' / 'false' / '' (a bare name means True), a lone value
  without '=' goes to the first option, unknown names raise."""
  arg = arg or ""
  if arg and "," not in arg and "=" not in arg:
    arg = f"{arg}=True" if arg in _ARG_SPEC else f"{next(iter(_ARG_SPEC))}={arg}"
  raw = {}
  for item in arg.split(","):
    if item:
      name, _, val = item.partition("=")
      raw[name] = val if "=" in item else "True"
  out = {}
  for name, default in _ARG_SPEC.items():
    val = raw.pop(name, None)
    if val is None:
      out[name] = default
    elif isinstance(default, bool):
      if val.lower() not in ("true", "false", ""):
        raise ValueError(f"{name}: expected a boolean, got '{val}'")
      out[name] = val.lower() == "true"
    else:
      out[name] = type(default)(val)
  if raw:
    raise ValueError(f"Unhandled config args remain: {raw}")
  return out


def get_config(arg: Optional[str] = None) -> Tuple[dict, TrainConfig]:
  """What configs/ae_i1k.py::get_config resolves for the training step, from the same option string: the `config.model`
  kwargs for `Model(**kw)` (ae_i1k.py:81-89) and a TrainConfig (ae_i1k.py:12-53,91-96).  Dataset, evaluator, logging and
  checkpoint entries of the reference config have no counterpart on this path."""
  a = parse_arg(arg)
  if a["latent_diffusion"]:
    if a["size"] != 256:
      raise AssertionError("Latent Diffusion only supports 256x256 images")     # ae_i1k.py:17
    space = (32, 32, 4)
  else:
    space = (a["size"], a["size"], 3)
  num_classes = 1000 if a["use_labels"] else None
  model = dict(num_classes=num_classes, variant=a["variant"], scan=a["scan"], adaln=a["adaln"], channels=space[-1],
               img_size=space[0], remat_policy="nothing_saveable")
  train = TrainConfig(batch_size=a["batch_size"], no_noise_prob=a["no_noise_prob"], mask_ratio=a["mask_ratio"],
                      mask_ratio_no_noise=a["mask_ratio_no_noise"], use_labels=a["use_labels"],
                      beta_schedule="linear" if a["latent_diffusion"] else "cosine", timesteps=1000, diffusion_space=space,
                      peak_lr=a["lr"], wd=a["wd"], betas=(0.9, a["beta2"]), clip_norm=1.0, total_epochs=a["epochs"],
                      ema_decay=0.0001 * (a["batch_size"] / 256) if a["use_labels"] else None)
  return model, train
