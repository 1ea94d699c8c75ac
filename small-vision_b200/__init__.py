"""small-vision_b200: B200-native training-step hot path of the small-vision UMD auto-encoder.

Public surface mirrors the reference's Python seams:
  Model / decode_variant            big_vision/models/ae.py:200-222
  create_gaussian_diffusion, q_sample   big_vision/gaussian_diffusion.py:32-98
  make_update_fn / update_fn        big_vision/trainers/train_ae.py:287-382
  infer_sharding                    big_vision/sharding.py:33-55
"""
__version__ = "0.1.0"
