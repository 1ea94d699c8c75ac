"""Few-shot ridge probe on `pre_logits` (big_vision/evaluators/fewshot_lsr.py).

Same names and meaning as the reference: `_precompute_cache(x, y, num_classes)` whitens the support set, appends the
bias feature and prepares the normal equations; `_eig_fewshot_acc_fn(cache, x_test, y_test, l2_reg)` solves them for one
penalty and returns the query accuracy; `Evaluator` selects `shots` support examples per class with the reference's
numpy permutations (fewshot_lsr.py:199-205) and reports the same metric names.  The reference factorises the Gram matrix
by `jnp.linalg.eigh` so that several penalties share one factorisation; the ridge solution is the same and is obtained
here by a Cholesky solve per penalty (`umd_fewshot_ridge_solve`, fp64 inside), with the Gram matrix cached instead.

Every array operation is a kernel of libumd_b200.so behind the C ABI (include/umd_b200.h, `umd_fewshot_*`); torch only
owns the device buffers.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib

BIAS_CONSTANT = 100.0   # fewshot_lsr.py:31


def _f32(x, dev="cuda"):
  x = torch.as_tensor(x)
  if not torch.cuda.is_available():
    raise lib.UmdError("the few-shot probe needs a CUDA device (no CPU fallback)")
  return x.to(device=dev, dtype=torch.float32).contiguous()


def _i32(y, dev="cuda"):
  return torch.as_tensor(y).to(device=dev, dtype=torch.int32).contiguous()


def matmul(A, B, *, trans_a=False, trans_b=False):
  """fp32 product of two row-major 2-D CUDA tensors (optionally transposed) through umd_fewshot_matmul."""
  assert A.is_cuda and B.is_cuda and A.dtype == torch.float32 and B.dtype == torch.float32
  assert A.dim() == 2 and B.dim() == 2 and A.is_contiguous() and B.is_contiguous()
  M, K = (A.shape[1], A.shape[0]) if trans_a else A.shape
  K2, N = (B.shape[1], B.shape[0]) if trans_b else B.shape
  assert K == K2, (A.shape, B.shape, trans_a, trans_b)
  a_rs, a_cs = (1, A.shape[1]) if trans_a else (A.shape[1], 1)
  b_rs, b_cs = (1, B.shape[1]) if trans_b else (B.shape[1], 1)
  out = torch.empty(M, N, dtype=torch.float32, device=A.device)
  L = lib.load()
  lib.check(L.umd_fewshot_matmul(lib.ptr(A), C.c_longlong(a_rs), C.c_longlong(a_cs), lib.ptr(B), C.c_longlong(b_rs),
                                 C.c_longlong(b_cs), lib.ptr(out), C.c_longlong(N), C.c_int(M), C.c_int(N), C.c_int(K),
                                 lib.current_stream()), "umd_fewshot_matmul")
  return out


def _whiten(x, mean, std):
  n, d = x.shape
  out = torch.empty(n, d + 1, dtype=torch.float32, device=x.device)
  lib.check(lib.load().umd_fewshot_whiten(lib.ptr(x), lib.ptr(mean), lib.ptr(std), C.c_int(n), C.c_int(d),
                                          C.c_float(BIAS_CONSTANT), lib.ptr(out), lib.current_stream()), "umd_fewshot_whiten")
  return out


def _precompute_cache(x, y, num_classes):
  """fewshot_lsr.py:43-97.  x: [N, D] features, y: [N] labels.  Returns the cache consumed by `_eig_fewshot_acc_fn`:
  {"mean" [1, D], "std" [1, D], "gram" (X^T X [D+1, D+1] if N >= D+1 else X X^T [N, N]), "rhs" (X^T Y or Y),
  "x" (the whitened support set, kept only when N < D+1), "num_classes"}."""
  L = lib.load()
  x, y = _f32(x), _i32(y, "cuda")
  n, d = x.shape
  dim = d + 1
  mean = torch.empty(1, d, dtype=torch.float32, device=x.device)
  std = torch.empty(1, d, dtype=torch.float32, device=x.device)
  lib.check(L.umd_fewshot_stats(lib.ptr(x), C.c_int(n), C.c_int(d), lib.ptr(mean), lib.ptr(std), lib.current_stream()),
            "umd_fewshot_stats")
  xw = _whiten(x, mean, std)
  cache = {"mean": mean, "std": std, "num_classes": int(num_classes)}
  if n >= dim:   # fewshot_lsr.py:80-83
    cache["gram"] = matmul(xw, xw, trans_a=True)
    sums = torch.empty((num_classes + 2) * dim, dtype=torch.float32, device=x.device)
    rhs = torch.empty(dim, num_classes, dtype=torch.float32, device=x.device)
    lib.check(L.umd_fewshot_xty(lib.ptr(xw), lib.ptr(y), C.c_int(n), C.c_int(dim), C.c_int(num_classes), lib.ptr(sums),
                                lib.ptr(rhs), lib.current_stream()), "umd_fewshot_xty")
    cache["rhs"], cache["x"] = rhs, None
  else:          # fewshot_lsr.py:84-88
    cache["gram"] = matmul(xw, xw, trans_b=True)
    tgt = torch.empty(n, num_classes, dtype=torch.float32, device=x.device)
    lib.check(L.umd_fewshot_targets(lib.ptr(y), C.c_int(n), C.c_int(num_classes), lib.ptr(tgt), lib.current_stream()),
              "umd_fewshot_targets")
    cache["rhs"], cache["x"] = tgt, xw
  return cache


def ridge_weights(cache, l2_reg):
  """w [D+1, num_classes] of fewshot_lsr.py:103-108 for one penalty."""
  L = lib.load()
  gram, rhs = cache["gram"], cache["rhs"]
  n, c = gram.shape[0], rhs.shape[1]
  L.umd_fewshot_solve_scratch_bytes.restype = C.c_size_t
  nbytes = int(L.umd_fewshot_solve_scratch_bytes(C.c_int(n), C.c_int(c)))
  scratch = torch.empty(nbytes, dtype=torch.uint8, device=gram.device)
  status = torch.zeros(1, dtype=torch.int32, device=gram.device)
  z = torch.empty_like(rhs)
  lib.check(L.umd_fewshot_ridge_solve(lib.ptr(gram), C.c_float(float(l2_reg)), lib.ptr(rhs), C.c_int(n), C.c_int(c),
                                      lib.ptr(z), lib.ptr(scratch), C.c_size_t(nbytes), lib.ptr(status),
                                      lib.current_stream()), "umd_fewshot_ridge_solve")
  st = int(status.item())
  if st != 0:
    raise lib.UmdError(f"few-shot ridge system is not positive definite at pivot {st - 1} (l2_reg = {l2_reg})")
  return z if cache["x"] is None else matmul(cache["x"], z, trans_a=True)


def _eig_fewshot_acc_fn(cache, x_test, y_test, l2_reg, return_preds=False):
  """fewshot_lsr.py:94-112: accuracy of the ridge regressor on (x_test, y_test) as a 0-d float32 CUDA tensor."""
  L = lib.load()
  x_test, y_test = _f32(x_test), _i32(y_test)
  w = ridge_weights(cache, l2_reg)
  scores = matmul(_whiten(x_test, cache["mean"], cache["std"]), w)
  n, c = scores.shape
  correct = torch.zeros(1, dtype=torch.int32, device=scores.device)
  preds = torch.empty(n, dtype=torch.int32, device=scores.device) if return_preds else None
  lib.check(L.umd_fewshot_accuracy(lib.ptr(scores), lib.ptr(y_test), C.c_int(n), C.c_int(c), lib.ptr(preds),
                                   lib.ptr(correct), lib.current_stream()), "umd_fewshot_accuracy")
  acc = correct.to(torch.float32)[0] / n
  return (acc, preds) if return_preds else acc


class Evaluator:
  """fewshot_lsr.py:115-240 without the TFDS input pipeline: `datasets` maps a name to in-memory
  (train_images, train_labels, test_images, test_labels) arrays (the reference resolves TFDS names and preprocessing
  strings instead; that stage is out of scope).  `predict_fn(train_state, batch)` is one of
  evaluators.make_predict_fn / create_noised_pred_fn; `representation_layer` is a key of its output dict."""

  def __init__(self, predict_fn, batch_size, representation_layer="pre_logits", datasets=None, shots=(100,),
               l2_reg=1024, num_seeds=3, display_first=(), num_classes=None):
    self.predict_fn = predict_fn
    self.batch_size = int(batch_size)
    self.representation_layer = representation_layer
    self.datasets = datasets or {}
    self.shots = tuple(shots)
    self.l2_reg = l2_reg
    self.num_seeds = num_seeds
    self.display_first = display_first
    self.num_classes = num_classes
    self._repr = {}

  def _get_repr(self, train_state, images, labels):
    """fewshot_lsr.py:176-191: representation of the whole split, batch by batch."""
    reps = []
    for i in range(0, len(images), self.batch_size):
      *_, out = self.predict_fn(train_state, {"image": torch.as_tensor(images[i:i + self.batch_size])})
      reps.append(out[self.representation_layer].float())
    return torch.cat(reps, 0), np.asarray(labels)

  def compute_fewshot_metrics(self, train_state, seed, name):
    tr_x, tr_y, te_x, te_y = self.datasets[name]
    if name not in self._repr:
      rtr, ltr = self._get_repr(train_state, tr_x, tr_y)
      rte, lte = self._get_repr(train_state, te_x, te_y)
      nc = self.num_classes or int(max(ltr.max(), lte.max())) + 1
      self._repr[name] = (rtr, ltr, rte, lte, nc)
    rtr, ltr, rte, lte, nc = self._repr[name]
    rng = np.random.default_rng(seed)                                    # fewshot_lsr.py:199-201
    class_indices = [rng.permutation(np.where(ltr == c)[0]) for c in range(nc)]
    results = {}
    for shots in self.shots:
      idx = np.concatenate([ind[:shots] for ind in class_indices], axis=0)
      sel = torch.as_tensor(idx, device=rtr.device)
      cache = _precompute_cache(rtr[sel], ltr[idx], nc)
      results[shots] = float(_eig_fewshot_acc_fn(cache, rte, lte, self.l2_reg))
    return results

  def run(self, train_state):
    """fewshot_lsr.py:226-236: yields (metric name, accuracy)."""
    self._repr = {}
    for seed in range(self.num_seeds):
      for name in self.datasets:
        for shots, v in self.compute_fewshot_metrics(train_state, seed, name).items():
          prefix = "a/" if (name, shots) in self.display_first else "z/"
          yield f"{prefix}{name}_{shots}shot-seed-{seed}", v
