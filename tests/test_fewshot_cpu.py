"""Few-shot ridge probe: the oracle's restatement (oracle/umd_oracle.py::fewshot_*) against tests/golden/fewshot_golden.pt,
the outputs of the reference's own `_precompute_cache` / `_eig_fewshot_acc_fn` (fewshot_lsr.py:43-112) executed over the
numpy-fp64 jax stand-in (tests/golden/make_fewshot_golden.py).  Also the closed-form identity the CUDA path relies on:
the eigendecomposition route equals the direct ridge solve."""
import os

import pytest
import torch

from oracle import umd_oracle as O
from tests import util as U
from tests.golden import make_fewshot_golden as FG

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fewshot_golden.pt"))


@pytest.mark.parametrize("name", sorted(FG.CASES))
def test_oracle_matches_reference_source(name):
  x, y, xt, yt, c, l2 = FG.make_case(name)
  want = GOLD["cases"][name]
  assert abs(float(x.double().sum()) - want["x_sum"]) <= 1e-6 * abs(want["x_sum"]), "inputs no longer regenerate"
  cache = O.fewshot_precompute_cache(x.double(), y, c)
  assert torch.allclose(cache["mean"].float(), want["mean"], rtol=1e-6, atol=1e-6)
  assert torch.allclose(cache["std"].float(), want["std"], rtol=1e-6, atol=1e-6)
  w = O.fewshot_weights(cache, l2)
  assert U.rel_l2(w, want["w"]) <= 1e-6
  acc, preds, _ = O.fewshot_acc(cache, xt.double(), yt, l2)
  assert torch.equal(preds.to(torch.int16), want["preds"])
  assert abs(acc - want["acc"]) <= 1e-12


@pytest.mark.parametrize("name", sorted(FG.CASES))
def test_eigh_route_equals_direct_ridge_solve(name):
  """(lhs diag(1/(eigs + l2)) rhs) == (X^T X + l2 I)^-1 X^T Y == X^T (X X^T + l2 I)^-1 Y  (fewshot_lsr.py:56-77)."""
  x, y, xt, yt, c, l2 = FG.make_case(name)
  cache = O.fewshot_precompute_cache(x.double(), y, c)
  xw = (x.double() - cache["mean"]) / cache["std"]
  xw = torch.cat([xw, torch.full((xw.shape[0], 1), O.FEWSHOT_BIAS_CONSTANT, dtype=torch.float64)], 1)
  yy = 2.0 * torch.nn.functional.one_hot(y.long(), c).double() - 1.0
  n, dim = xw.shape
  a = torch.linalg.solve(xw.T @ xw + l2 * torch.eye(dim, dtype=torch.float64), xw.T @ yy)
  b = xw.T @ torch.linalg.solve(xw @ xw.T + l2 * torch.eye(n, dtype=torch.float64), yy)
  w = O.fewshot_weights(cache, l2)
  assert U.rel_l2(a, w) <= 1e-8 and U.rel_l2(b, w) <= 1e-8
