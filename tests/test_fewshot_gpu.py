"""Few-shot ridge probe on the GPU (small-vision_b200/fewshot.py -> umd_fewshot_* through the C ABI) against
tests/golden/fewshot_golden.pt — outputs of the reference's own fewshot_lsr.py functions executed over the numpy-fp64 jax
stand-in — and against the oracle at the ImageNet-probe shape.  Tolerances: the CUDA path accumulates the Gram matrix
in fp32 (the reference computes in fp32 throughout), so weights agree to 1e-3 relative and every prediction whose
reference score margin exceeds 1e-3 must be identical."""
import os

import pytest
import torch

from tests import util as U
from tests.golden import make_fewshot_golden as FG

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fewshot_golden.pt"))


def test_matmul_kernel_all_layouts():
  from small_vision_b200 import fewshot as FS
  g = torch.Generator().manual_seed(0)
  for (m, n, k) in ((1, 1, 1), (130, 70, 33), (257, 129, 1000), (769, 769, 5000)):
    a, b = torch.randn(m, k, generator=g), torch.randn(k, n, generator=g)
    ref = a.double() @ b.double()
    for ta in (False, True):
      for tb in (False, True):
        A = (a.T if ta else a).contiguous().to(DEV)
        B = (b.T if tb else b).contiguous().to(DEV)
        got = FS.matmul(A, B, trans_a=ta, trans_b=tb)
        assert U.rel_l2(got.cpu(), ref) <= 2e-6, (m, n, k, ta, tb)


@pytest.mark.parametrize("name", sorted(FG.CASES))
def test_probe_matches_reference_source(name):
  from small_vision_b200 import fewshot as FS
  x, y, xt, yt, c, l2 = FG.make_case(name)
  want = GOLD["cases"][name]
  cache = FS._precompute_cache(x.to(DEV), y.to(DEV), c)
  assert torch.allclose(cache["mean"].cpu(), want["mean"], rtol=1e-5, atol=1e-5)
  assert torch.allclose(cache["std"].cpu(), want["std"], rtol=1e-5, atol=1e-5)
  assert (cache["x"] is None) == (x.shape[0] >= x.shape[1] + 1)
  w = FS.ridge_weights(cache, l2)
  assert U.rel_l2(w.cpu(), want["w"]) <= 1e-3, U.rel_l2(w.cpu(), want["w"])
  acc, preds = FS._eig_fewshot_acc_fn(cache, xt.to(DEV), yt.to(DEV), l2, return_preds=True)
  preds = preds.cpu().to(torch.int16)
  clear = want["margin"] > 1e-3
  assert torch.equal(preds[clear], want["preds"][clear])
  n_unclear = int((~clear).sum())
  assert abs(float(acc) - want["acc"]) <= n_unclear / len(preds) + 1e-7
  # the counted accuracy is the accuracy of the returned predictions
  assert abs(float(acc) - float((preds.long() == yt.long()).float().mean())) <= 1e-7


def test_penalty_sweep_reuses_the_cache():
  """fewshot_lsr.py:56-77: one cache serves any l2; larger penalties shrink the weights monotonically."""
  from small_vision_b200 import fewshot as FS
  x, y, xt, yt, c, _ = FG.make_case("tall_n_ge_d")
  cache = FS._precompute_cache(x.to(DEV), y.to(DEV), c)
  norms = [float(FS.ridge_weights(cache, l2).norm()) for l2 in (1.0, 32.0, 1024.0, 32768.0)]
  assert all(a > b for a, b in zip(norms, norms[1:])), norms


def test_imagenet_probe_shape_against_oracle():
  """10 shots x 1000 classes at width 768 (configs/eval_ae_i1k.py:122): N = 10 000 >= 769, C = 1000."""
  from oracle import umd_oracle as O
  from small_vision_b200 import fewshot as FS
  g = torch.Generator().manual_seed(9)
  c, d, shots, nt = 1000, 768, 10, 4000
  centres = torch.randn(c, d, generator=g) * 0.25
  y = torch.arange(c).repeat_interleave(shots)
  x = centres[y] + torch.randn(c * shots, d, generator=g)
  yt = torch.randint(0, c, (nt,), generator=g)
  xt = centres[yt] + torch.randn(nt, d, generator=g)
  cache = FS._precompute_cache(x.to(DEV), y.to(DEV), c)
  acc, preds = FS._eig_fewshot_acc_fn(cache, xt.to(DEV), yt.to(DEV), 1024.0, return_preds=True)
  ocache = O.fewshot_precompute_cache(x.double(), y, c)
  oacc, opreds, oscores = O.fewshot_acc(ocache, xt.double(), yt, 1024.0)
  w = FS.ridge_weights(cache, 1024.0)
  assert U.rel_l2(w.cpu(), O.fewshot_weights(ocache, 1024.0)) <= 1e-3
  top2 = oscores.topk(2, dim=1).values
  clear = (top2[:, 0] - top2[:, 1]) > 1e-3
  assert torch.equal(preds.cpu().long()[clear], opreds[clear])
  assert abs(float(acc) - oacc) <= float((~clear).sum()) / nt + 1e-7


def test_evaluator_reports_reference_metric_names():
  """Evaluator.run over an in-memory dataset through the model's predict_fn (fewshot_lsr.py:193-236)."""
  from small_vision_b200 import fewshot as FS
  from small_vision_b200.evaluators import make_predict_fn
  model, _ = U.make_models("S/4", adaln=True, depth=1, dec_depth=1)
  params = U.perturb_init(model, 0, DEV)
  g = torch.Generator().manual_seed(3)
  protos = torch.rand(4, 64, 64, 3, generator=g) * 2 - 1
  def split(m):
    y = torch.arange(m) % 4
    return (protos[y] * 0.7 + 0.3 * (torch.rand(m, 64, 64, 3, generator=g) * 2 - 1)), y.numpy()
  tr_x, tr_y = split(48)
  te_x, te_y = split(32)
  ev = FS.Evaluator(make_predict_fn(model), batch_size=16, datasets={"toy": (tr_x, tr_y, te_x, te_y)}, shots=(2, 8),
                    l2_reg=1024, num_seeds=2)
  res = dict(ev.run({"params": params}))
  assert sorted(res) == sorted(f"z/toy_{s}shot-seed-{k}" for s in (2, 8) for k in (0, 1))
  assert all(0.0 <= v <= 1.0 for v in res.values())
  assert res["z/toy_8shot-seed-0"] >= 0.75, res     # four well-separated prototypes
