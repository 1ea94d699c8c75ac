"""Generates tests/golden/fewshot_golden.pt by executing the reference's own `_precompute_cache` and
`_eig_fewshot_acc_fn` (big_vision/evaluators/fewshot_lsr.py:43-112, unmodified, imported from /root/reference) over the
numpy-fp64 jax stand-in of tests/golden/refshim — `jnp.linalg.eigh` is numpy's LAPACK eigh there, an implementation
independent of both the oracle's (torch) and the CUDA path's (Cholesky).  Inputs are regenerated from seeds by the tests.

  python tests/golden/make_fewshot_golden.py       (build container only: needs /root/reference)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("UMD_REFERENCE_ROOT", "/root/reference")

# name: (support size N, feature width D, classes, query size, l2_reg, seed, class-centre spread)
CASES = {
    "tall_n_ge_d": (600, 96, 10, 400, 1024.0, 1, 0.35),     # N >= D + 1: x^T x branch
    "wide_d_gt_n": (40, 96, 10, 300, 1024.0, 2, 0.35),      # D + 1 > N: x x^T branch
    "square_edge": (97, 96, 5, 200, 16.0, 3, 0.35),         # N == D + 1 exactly (first branch), small penalty
    "width_768": (1000, 768, 10, 500, 1024.0, 4, 0.09),     # the B/4 representation width, 100 shots x 10 classes
}


def make_case(name):
  """Class-structured features (so that the accuracy is neither 0 nor 1) drawn from a seeded torch generator."""
  n, d, c, nt, l2, seed, spread = CASES[name]
  g = torch.Generator().manual_seed(seed)
  centres = torch.randn(c, d, generator=g) * spread
  scale = torch.rand(d, generator=g) * 3 + 0.2
  offset = torch.randn(d, generator=g) * 2

  def draw(m):
    y = torch.arange(m) % c
    y = y[torch.randperm(m, generator=g)]
    x = (centres[y] + torch.randn(m, d, generator=g)) * scale + offset
    return x, y.to(torch.int32)
  x, y = draw(n)
  xt, yt = draw(nt)
  return x, y, xt, yt, c, l2


def load_reference():
  sys.path.insert(0, os.path.join(HERE, "refshim"))
  sys.path.insert(0, REF)
  for name in ("big_vision.utils", "big_vision.datasets", "big_vision.datasets.core", "big_vision.input_pipeline",
               "big_vision.pp", "big_vision.pp.builder"):   # module-level imports of the Evaluator class (TFDS / tf.data)
    sys.modules[name] = types.ModuleType(name)
  from big_vision.evaluators import fewshot_lsr
  assert fewshot_lsr.__file__.startswith(REF)
  return fewshot_lsr


def main():
  fs = load_reference()
  gold = {"provenance": "big_vision/evaluators/fewshot_lsr.py executed over tests/golden/refshim (numpy fp64)", "cases": {}}
  for name in CASES:
    x, y, xt, yt, c, l2 = make_case(name)
    J = lambda a: np.asarray(a.double().numpy()) if a.dtype.is_floating_point else np.asarray(a.numpy())
    cache = fs._precompute_cache(J(x), J(y), c)
    acc = float(fs._eig_fewshot_acc_fn(cache, J(xt), J(yt), l2))
    # the weights and predictions the accuracy came from (same expressions as fewshot_lsr.py:103-111)
    w = (np.asarray(cache["lhs"]) * (1.0 / (np.asarray(cache["eigs"]) + l2)).reshape(1, -1)) @ np.asarray(cache["rhs"])
    xw = np.pad((J(xt) - np.asarray(cache["mean"])) / np.asarray(cache["std"]), ((0, 0), (0, 1)),
                constant_values=fs.BIAS_CONSTANT)
    scores = xw @ w
    preds = scores.argmax(1)
    assert abs(float((preds == J(yt)).mean()) - acc) < 1e-12
    top2 = np.sort(scores, axis=1)[:, -2:]
    gold["cases"][name] = {"acc": acc, "w": torch.from_numpy(w).float(), "preds": torch.from_numpy(preds).to(torch.int16),
                           "margin": torch.from_numpy(top2[:, 1] - top2[:, 0]).float(),
                           "mean": torch.from_numpy(np.asarray(cache["mean"])).float(),
                           "std": torch.from_numpy(np.asarray(cache["std"])).float(),
                           "x_sum": float(x.double().sum())}
    print(name, "acc", acc, "w norm", float(np.linalg.norm(w)))
  path = os.path.join(HERE, "fewshot_golden.pt")
  torch.save(gold, path)
  print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
  main()
