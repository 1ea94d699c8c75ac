"""small-vision_b200: B200-native training-step hot path of the small-vision UMD auto-encoder.

Public surface mirrors the reference's Python seams:
  Model / decode_variant            big_vision/models/ae.py:200-222
  create_gaussian_diffusion, q_sample   big_vision/gaussian_diffusion.py:32-98
  ddim_sample, ddim_sample_loop     big_vision/gaussian_diffusion.py:134-284
  make_update_fn / update_fn        big_vision/trainers/train_ae.py:287-382
  infer_sharding                    big_vision/sharding.py:33-55
"""
__version__ = "0.1.0"

from .config import ModelConfig, TrainConfig, decode_variant, make_model_config  # noqa: E402,F401
from .sharding import infer_sharding, Mesh, PartitionSpec  # noqa: E402,F401


def __getattr__(name):
  # torch / CUDA dependent pieces are imported lazily so that host-only tooling can import the package
  if name in ("Model", "ViTAE", "mask_argsort"):
    from . import model as _m
    return getattr(_m, name)
  if name in ("create_gaussian_diffusion", "q_sample", "get_beta_schedule", "ddim_sample", "ddim_sample_loop",
              "create_apply_fn", "reference_timesteps"):
    from . import diffusion as _d
    return getattr(_d, name)
  if name in ("make_update_fn", "create_train_state"):
    from . import train as _t
    return getattr(_t, name)
  raise AttributeError(name)
