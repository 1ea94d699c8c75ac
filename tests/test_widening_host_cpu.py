"""Host-side logic of the §8f rank 3 / rank 4 components that needs no GPU: crop-window sampling constraints
(pp/ops_image.py:226-233), the support-set selection of the few-shot evaluator (fewshot_lsr.py:199-205), and the loud
failure of both modules without a CUDA device / library."""
import numpy as np
import pytest
import torch


def test_inception_boxes_respect_area_and_ratio_ranges():
  from small_vision_b200 import pp
  H, W = 96, 128
  b = pp.sample_inception_boxes(500, H, W, area_min=5, area_max=100, ratio_min=0.75, ratio_max=1.33, seed=3)
  assert b.dtype == np.int32 and b.shape == (500, 4)
  y0, x0, h, w = b.T
  assert (y0 >= 0).all() and (x0 >= 0).all() and (h > 0).all() and (w > 0).all()
  assert (y0 + h <= H).all() and (x0 + w <= W).all()
  whole = (h == H) & (w == W)
  area = (h * w) / (H * W)
  ratio = w / h
  ok = ~whole
  assert (area[ok] >= 0.05 - 0.02).all() and (area[ok] <= 1.0).all()      # integer rounding of h, w
  assert (ratio[ok] >= 0.75 - 0.05).all() and (ratio[ok] <= 1.33 + 0.05).all()
  assert np.array_equal(b, pp.sample_inception_boxes(500, H, W, seed=3))  # seeded
  assert not np.array_equal(b, pp.sample_inception_boxes(500, H, W, seed=4))


def test_fewshot_support_selection_follows_the_reference_rule():
  """Per seed: one numpy permutation of each class's indices in class order, first `shots` of each (fewshot_lsr.py:199-205);
  reproduced here independently and compared with what Evaluator feeds to _precompute_cache."""
  from small_vision_b200 import fewshot as FS
  labels = np.array([0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 0, 1])
  feats = torch.arange(len(labels), dtype=torch.float32)[:, None].repeat(1, 4)
  seen = []

  class Ev(FS.Evaluator):
    def _get_repr(self, train_state, images, labels_):
      return feats, np.asarray(labels_)
  orig_pc, orig_acc = FS._precompute_cache, FS._eig_fewshot_acc_fn
  FS._precompute_cache = lambda x, y, nc: seen.append((x[:, 0].tolist(), list(y), nc)) or {}
  FS._eig_fewshot_acc_fn = lambda cache, xt, yt, l2: 0.5
  try:
    ev = Ev(None, 4, datasets={"toy": (None, labels, None, labels)}, shots=(1, 2), num_seeds=1)
    res = dict(ev.run({}))
  finally:
    FS._precompute_cache, FS._eig_fewshot_acc_fn = orig_pc, orig_acc
  assert sorted(res) == ["z/toy_1shot-seed-0", "z/toy_2shot-seed-0"]
  rng = np.random.default_rng(0)
  perms = [rng.permutation(np.where(labels == c)[0]) for c in range(3)]
  for (xs, ys, nc), shots in zip(seen, (1, 2)):
    want = np.concatenate([p[:shots] for p in perms])
    assert nc == 3 and xs == [float(i) for i in want] and ys == [int(labels[i]) for i in want]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
  from small_vision_b200 import fewshot as FS, lib, pp
  with pytest.raises(lib.UmdError):
    pp.augment(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))
  with pytest.raises(lib.UmdError):
    FS._precompute_cache(torch.zeros(4, 3), torch.zeros(4, dtype=torch.int32), 2)
