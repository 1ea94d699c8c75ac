"""Oracle restatement of the post-decode input stage (oracle/umd_oracle.py::preprocess_train; TensorFlow, which owns this
arithmetic in the reference, is absent — parity unpinned): known answers derivable from the half-pixel-centre
definition, and agreement with torch's independent bilinear implementation (same sampling convention)."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import umd_oracle as O


def _img(n, H, W, C, seed):
  return np.random.default_rng(seed).integers(0, 256, (n, H, W, C), dtype=np.uint8)


def test_identity_window_and_value_range_endpoints():
  x = _img(3, 16, 16, 3, 0)
  x[0, 0, 0] = (0, 255, 128)
  out, mid = O.preprocess_train(x)
  assert np.array_equal(mid, x)
  assert out[0, 0, 0, 0] == -1.0 and out[0, 0, 0, 1] == 1.0
  assert out[0, 0, 0, 2] == np.float32(-1) + (np.float32(128) / np.float32(255)) * np.float32(2)   # pp/ops_general.py:56-57


def test_halving_is_the_2x2_box_mean_truncated():
  """scale 2: source coordinate 2i + 0.5 => equal weights on the 2 x 2 block (resize then tf.cast truncation)."""
  x = _img(2, 32, 48, 3, 1)
  _, mid = O.preprocess_train(x, boxes=[(0, 0, 32, 32), (0, 16, 32, 32)], size=16)
  for i, x0 in enumerate((0, 16)):
    c = x[i, :, x0:x0 + 32].astype(np.float32)
    top = c[0::2, 0::2] + (c[0::2, 1::2] - c[0::2, 0::2]) * np.float32(0.5)
    bot = c[1::2, 0::2] + (c[1::2, 1::2] - c[1::2, 0::2]) * np.float32(0.5)
    want = (top + (bot - top) * np.float32(0.5)).astype(np.uint8)
    assert np.array_equal(mid[i], want)


def test_flip_is_a_mirror_and_an_involution():
  x = _img(2, 20, 24, 3, 2)
  _, a = O.preprocess_train(x, flips=[True, False])
  assert np.array_equal(a[0], x[0, :, ::-1]) and np.array_equal(a[1], x[1])


def test_matches_torch_bilinear_up_to_the_truncation_boundary():
  x = _img(4, 45, 61, 3, 3)
  boxes = [(0, 0, 45, 61), (3, 7, 30, 40), (10, 1, 9, 55), (44, 60, 1, 1)]
  _, mid = O.preprocess_train(x, boxes=boxes, size=32)
  for i, (y0, x0, h, w) in enumerate(boxes):
    crop = torch.from_numpy(x[i, y0:y0 + h, x0:x0 + w].astype(np.float32)).permute(2, 0, 1)[None]
    ref = F.interpolate(crop.double(), size=(32, 32), mode="bilinear", align_corners=False, antialias=False)[0]
    ref = ref.permute(1, 2, 0).numpy()
    got = mid[i].astype(np.float64)
    # truncation: got == floor(ref) except where ref sits within float32 round-off of an integer
    assert np.all((got <= ref + 1e-3) & (got >= ref - 1 - 1e-3))
    assert np.mean(got == np.floor(ref)) > 0.999


def test_value_range_matches_reference_source():
  """oracle value-range stage == the reference's own get_value_range (pp/ops_general.py:30-62) run over a numpy stand-in
  for its five tf names (tests/golden/make_pp_golden.py), bit for bit, on every uint8 value."""
  import os
  from tests.golden import make_pp_golden as PG
  gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_pp_golden.npz"))
  x = PG.all_bytes()
  for i, (vmin, vmax, in_min, in_max, clip) in enumerate(PG.SETTINGS):
    out, mid = O.preprocess_train(x, vmin=vmin, vmax=vmax, in_min=in_min, in_max=in_max, clip_values=clip)
    assert np.array_equal(mid, x)
    assert np.array_equal(out, gold[f"case{i}"]), PG.SETTINGS[i]
