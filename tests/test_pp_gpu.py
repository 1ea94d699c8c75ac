"""Fused GPU input stage (small-vision_b200/pp.py -> umd_augment_u8) against the oracle's float32 restatement of the
reference's preprocessing string after JPEG decoding (configs/ae_i1k.py:64-69).  Byte / float outputs must be
BIT-EXACT: both evaluate the same IEEE single-precision operations in the same order."""
import numpy as np
import pytest
import torch

from oracle import umd_oracle as O

pytestmark = pytest.mark.gpu


def _images(n, H, W, C, seed):
  g = torch.Generator().manual_seed(seed)
  return torch.randint(0, 256, (n, H, W, C), generator=g, dtype=torch.uint8)


def test_value_range_alone_is_exact():
  from small_vision_b200 import pp
  x = _images(5, 64, 64, 3, 0)
  x[0, 0, 0, 0], x[0, 0, 0, 1] = 0, 255
  got = pp.get_value_range(-1, 1)({"image": x})["image"].cpu().numpy()
  want, _ = O.preprocess_train(x.numpy())
  assert got.dtype == np.float32 and np.array_equal(got, want)
  assert got.min() == -1.0 and got.max() == 1.0


@pytest.mark.parametrize("H,W,S,C", [(64, 64, 64, 3), (96, 80, 64, 3), (37, 53, 32, 1), (256, 256, 64, 3), (40, 40, 64, 4),
                                     (50, 70, (24, 40), 3)])
def test_crop_resize_flip_value_range_bit_exact(H, W, S, C):
  from small_vision_b200 import pp
  n = 9
  x = _images(n, H, W, C, H * 1000 + W)
  boxes = pp.sample_inception_boxes(n, H, W, area_min=5, area_max=100, seed=H + W)
  boxes[0] = (0, 0, H, W)                 # whole image
  boxes[1] = (H - 1, W - 1, 1, 1)         # single pixel in the corner
  boxes[2] = (0, 0, 1, W)                 # one row
  flips = np.arange(n) % 2 == 1
  got, got_u8 = pp.augment(x, boxes=boxes, flips=flips, size=S, return_uint8=True)
  want, want_u8 = O.preprocess_train(x.numpy(), boxes=boxes, flips=flips, size=S)
  assert np.array_equal(got_u8.cpu().numpy(), want_u8)
  assert np.array_equal(got.cpu().numpy(), want)


def test_reference_op_names_compose_to_the_fused_launch():
  """decode_jpeg_and_inception_crop | flip_lr | value_range one op at a time == the fused make_train_preprocess."""
  from small_vision_b200 import pp
  n, H, W = 6, 72, 88
  x = _images(n, H, W, 3, 5)
  boxes = pp.sample_inception_boxes(n, H, W, seed=11)
  flips = np.array([1, 0, 1, 1, 0, 0], dtype=bool)
  data = {"image": x, "label": torch.arange(n), "_boxes": boxes, "_flips": flips}
  fused = pp.make_train_preprocess(64)(data)
  step = pp.get_decode_jpeg_and_inception_crop(size=64)(data)
  step = pp.get_random_flip_lr()(step)
  step = pp.get_value_range(-1, 1)(step)
  assert torch.equal(fused["image"], step["image"])
  assert sorted(fused) == ["image", "label"]
  # properties: identity window at the native size is the plain rescale; flipping twice restores the image
  same = pp.augment(x, size=None).cpu()
  assert torch.equal(same, x.float() / 255.0 * 2.0 - 1.0) or torch.allclose(same, x.float() / 255.0 * 2.0 - 1.0, atol=1e-7)
  _, once = pp.augment(x, flips=np.ones(n, bool), return_uint8=True)
  _, twice = pp.augment(once, flips=np.ones(n, bool), return_uint8=True)
  assert torch.equal(twice.cpu(), x)


def test_out_of_range_window_raises():
  from small_vision_b200 import pp
  x = _images(1, 16, 16, 3, 0)
  with pytest.raises(ValueError):
    pp.augment(x, boxes=np.array([[8, 8, 16, 4]]), size=8)
