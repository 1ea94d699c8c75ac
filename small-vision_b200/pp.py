"""Training input stage on the GPU: what the reference's tf.data preprocessing string
`decode_jpeg_and_inception_crop(size)|flip_lr|value_range(-1, 1)` (configs/ae_i1k.py:64-69) does after JPEG decoding,
as ONE fused kernel over a uint8 batch (`umd_augment_u8`), plus the reference's op factories by name
(pp/ops_general.py:30-62, pp/ops_image.py:56-85,163-242,306-314) for callers that compose them one at a time.

The reference's ops map a per-example dict of tf tensors; here they map a batch dict {"image": uint8 or float CUDA
tensor [n, H, W, C], ...}.  JPEG decoding and the TFDS / tf.data plumbing stay out of scope.  Random draws (crop
windows, flips) are explicit arguments — `sample_inception_boxes` draws windows with the same constraints as
tf.image.sample_distorted_bounding_box but not its bit stream (like every other draw of the path).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib


def augment(images, *, boxes=None, flips=None, size=None, vmin=-1.0, vmax=1.0, in_min=0.0, in_max=255.0,
            clip_values=False, return_uint8=False):
  """crop -> bilinear resize (uint8 round trip) -> flip -> value_range in one launch.
  images: uint8 [n, Hs, Ws, C]; size: None (source size), int or (h, w); boxes: int [n, 4] = (y0, x0, h, w) or None; flips: bool [n] or None."""
  if not torch.cuda.is_available():
    raise lib.UmdError("the input stage needs a CUDA device (no CPU fallback)")
  images = torch.as_tensor(images)
  assert images.dtype == torch.uint8 and images.dim() == 4, (images.dtype, images.shape)
  images = images.cuda().contiguous()
  n, Hs, Ws, Cc = images.shape
  oh, ow = (Hs, Ws) if size is None else ((int(size), int(size)) if np.isscalar(size) else (int(size[0]), int(size[1])))
  b = f = None
  if boxes is not None:
    bh = torch.as_tensor(boxes).to(torch.int64).cpu().reshape(n, 4)
    ok = (bh[:, 0] >= 0) & (bh[:, 1] >= 0) & (bh[:, 2] > 0) & (bh[:, 3] > 0) & \
         (bh[:, 0] + bh[:, 2] <= Hs) & (bh[:, 1] + bh[:, 3] <= Ws)
    if not bool(ok.all()):
      raise ValueError("crop window outside the image")
    b = bh.to(torch.int32).cuda().contiguous()
  if flips is not None:
    f = torch.as_tensor(flips).reshape(n).to(torch.uint8).cuda().contiguous()
  out = torch.empty(n, oh, ow, Cc, dtype=torch.float32, device=images.device)
  u8 = torch.empty(n, oh, ow, Cc, dtype=torch.uint8, device=images.device) if return_uint8 else None
  lib.check(lib.load().umd_augment_u8(lib.ptr(images), C.c_int(n), C.c_int(Hs), C.c_int(Ws), C.c_int(Cc), lib.ptr(b),
                                      lib.ptr(f), C.c_int(oh), C.c_int(ow), C.c_float(in_min), C.c_float(in_max), C.c_double(vmin),
                                      C.c_double(vmax), C.c_int(int(clip_values)), lib.ptr(out), lib.ptr(u8),
                                      lib.current_stream()), "umd_augment_u8")
  return (out, u8) if return_uint8 else out


def sample_inception_boxes(n, height, width, *, area_min=5, area_max=100, ratio_min=0.75, ratio_max=1.33, seed=0,
                           max_attempts=100):
  """Crop windows with the constraints of pp/ops_image.py:226-233 (area fraction in [area_min, area_max] %, aspect ratio
  in [ratio_min, ratio_max], whole image when no attempt fits).  Host-side numpy; int32 [n, 4] = (y0, x0, h, w)."""
  rng = np.random.default_rng(seed)
  out = np.empty((n, 4), dtype=np.int32)
  for i in range(n):
    box = (0, 0, height, width)
    for _ in range(max_attempts):
      ratio = rng.uniform(ratio_min, ratio_max)
      area = rng.uniform(area_min / 100.0, area_max / 100.0) * height * width
      h = int(round(np.sqrt(area / ratio)))
      w = int(round(h * ratio))
      if 0 < h <= height and 0 < w <= width:
        box = (int(rng.integers(0, height - h + 1)), int(rng.integers(0, width - w + 1)), h, w)
        break
    out[i] = box
  return out


# ---- the reference's op factories by name; each returns fn(batch dict) -> batch dict -------------------------------
def get_value_range(vmin=-1, vmax=1, in_min=0, in_max=255.0, clip_values=False):
  """pp/ops_general.py:30-62 on a uint8 batch."""
  def _value_range(data):
    return {**data, "image": augment(data["image"], vmin=vmin, vmax=vmax, in_min=in_min, in_max=in_max,
                                     clip_values=clip_values)}
  return _value_range


class _Draws:
  """Per-op random stream: seeded once (optionally per data-parallel rank) and advanced on every call, so that successive
  batches get fresh crop windows / flips like the reference's per-example tf.random ops; data["_seed"] (an explicit
  per-batch seed) still overrides it for reproducible runs."""

  def __init__(self, seed=0, rank=0):
    self.ss = np.random.SeedSequence([int(seed), int(rank)])
    self.calls = 0

  def next_seed(self, data):
    if "_seed" in data:
      return int(data["_seed"])
    self.calls += 1
    return int(np.random.SeedSequence([*self.ss.entropy, self.calls]).generate_state(1, dtype=np.uint64)[0] >> 1)


def get_decode_jpeg_and_inception_crop(size=None, area_min=5, area_max=100, ratio_min=0.75, ratio_max=1.33,
                                       method="bilinear", antialias=False, seed=0, rank=0):
  """pp/ops_image.py:197-242 on already decoded uint8 images: returns the resized crop as uint8.  The windows come from
  data["_boxes"] when present, else from sample_inception_boxes seeded by data.get("_seed", 0)."""
  assert method == "bilinear" and not antialias, "only the reference recipe's resize (bilinear, antialias=False) is fused"
  draws = _Draws(seed, rank)

  def _inception_crop(data):
    img = torch.as_tensor(data["image"])
    n, H, W, _ = img.shape
    boxes = data.get("_boxes")
    if boxes is None:
      boxes = sample_inception_boxes(n, H, W, area_min=area_min, area_max=area_max, ratio_min=ratio_min,
                                     ratio_max=ratio_max, seed=draws.next_seed(data))
    _, u8 = augment(img, boxes=boxes, size=size, return_uint8=True)
    return {**data, "image": u8}
  return _inception_crop


def get_random_flip_lr(seed=0, rank=0):
  """pp/ops_image.py:306-314: flips each image with probability 1/2 (data["_flips"] supplies the draws)."""
  draws = _Draws(seed + 1, rank)

  def _random_flip_lr_pp(data):
    img = torch.as_tensor(data["image"])
    flips = data.get("_flips")
    if flips is None:
      flips = np.random.default_rng(draws.next_seed(data) + 1).random(img.shape[0]) < 0.5
    _, u8 = augment(img, flips=flips, return_uint8=True)
    return {**data, "image": u8}
  return _random_flip_lr_pp


def make_train_preprocess(size, area_min=5, area_max=100, vmin=-1, vmax=1, seed=0, rank=0):
  """The whole training string of configs/ae_i1k.py:64-69 after decoding, as a single launch."""
  draws = _Draws(seed, rank)

  def _pp(data):
    img = torch.as_tensor(data["image"])
    n, H, W, _ = img.shape
    boxes = data.get("_boxes")
    sd = draws.next_seed(data) if (boxes is None or data.get("_flips") is None) else 0
    if boxes is None:
      boxes = sample_inception_boxes(n, H, W, area_min=area_min, area_max=area_max, seed=sd)
    flips = data.get("_flips")
    if flips is None:
      flips = np.random.default_rng(sd + 1).random(n) < 0.5
    out = {"image": augment(img, boxes=boxes, flips=flips, size=size, vmin=vmin, vmax=vmax)}
    if "label" in data:
      out["label"] = data["label"]       # keep("image", "label")
    return out
  return _pp
