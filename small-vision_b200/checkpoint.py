"""Checkpoint import / export in the reference's flat `.npz` layout.

big_vision stores a pytree as one array per leaf, named by the '/'-joined path of the leaf with dict keys visited in
sorted order (`_traverse_with_names`, big_vision/utils.py:650-673) and rebuilds it with `recover_tree`
(utils.py:857-884); `load_params` (utils.py:239-290) accepts a file holding the whole train state ("params/..."),
a Flax-optimizer state ("opt/target/...") or the bare parameter tree, and an optional ":sub/key" suffix on the path.
The functions below keep those names and behaviours, so a `.npz` written by the reference's tooling loads into
`Model.apply` / `create_train_state` unchanged, and one written here loads there.  (The tensorstore directory format
of utils.py:886-1016 needs the `tensorstore` package, which this image does not have; it is out of scope.)

Host-side only: numpy in, numpy out; no CUDA, no torch requirement.
"""
from __future__ import annotations

import collections
import io
import os
import re
from collections.abc import Mapping

import numpy as np


def _to_numpy(v):
  if hasattr(v, "detach"):  # torch tensor (any device, incl. bf16 optimiser moments)
    v = v.detach()
    if str(v.dtype) == "torch.bfloat16":
      v = v.float()
    return v.cpu().numpy()
  return np.asarray(v)


def traverse_with_names(tree, prefix=()):
  """Leaf names as the reference's `_traverse_with_names` builds them (utils.py:650-673): path components joined by
  '/', mapping keys visited in sorted order, sequence items by index, `None` sub-trees skipped."""
  if tree is None:
    return
  if isinstance(tree, Mapping):
    children = [(str(k), tree[k]) for k in sorted(tree.keys())]
  elif isinstance(tree, (list, tuple)):
    children = [(str(i), v) for i, v in enumerate(tree)]
  else:
    yield "/".join(prefix), tree
    return
  for name, child in children:
    yield from traverse_with_names(child, prefix + (name,))


def tree_flatten_with_names(tree):
  """utils.py:676-706 (order = sorted-key depth-first, which is also jax's dict flattening order)."""
  return list(traverse_with_names(tree))


def recover_tree(keys, values):
  """Inverse of the flat naming (utils.py:857-884): nested dicts from '/'-separated names."""
  tree = {}
  for name, value in zip(keys, values):
    *parents, leaf = name.split("/")
    node = tree
    for part in parents:
      node = node.setdefault(part, {})
      if not isinstance(node, dict):
        raise ValueError(f"checkpoint name '{name}' descends through the leaf '{part}'")
    node[leaf] = value
  return tree


def tree_get(tree, name):
  """utils.py tree_get: sub-tree (or leaf) at a '/'-separated path."""
  node = tree
  for part in name.split("/"):
    if not isinstance(node, Mapping) or part not in node:
      raise KeyError(f"'{name}' not found in checkpoint (stopped at '{part}')")
    node = node[part]
  return node


def save_checkpoint_np(path, checkpoint):
  """One array per leaf under its '/'-joined name (the layout npload / load_checkpoint_np read back, utils.py:200-236).
  Written to a temporary file first and renamed, so an interrupted save never leaves a truncated checkpoint."""
  names_and_vals = tree_flatten_with_names(checkpoint)
  buf = io.BytesIO()
  np.savez(buf, **{k: _to_numpy(v) for k, v in names_and_vals})
  tmp = f"{path}.tmp-{os.getpid()}"
  with open(tmp, "wb") as f:
    f.write(buf.getvalue())
  os.replace(tmp, path)
  return [k for k, _ in names_and_vals]


def npload(fname):
  """utils.py:200-215."""
  loaded = np.load(fname, allow_pickle=False)
  if isinstance(loaded, np.ndarray):
    return loaded
  return dict(loaded)


def load_checkpoint_np(npz):
  """utils.py:218-236."""
  if isinstance(npz, (str, os.PathLike)):
    npz = npload(npz)
  keys, values = zip(*list(npz.items()))
  return recover_tree(keys, values)


def load_params(ckpt):
  """utils.py:239-290 for `.npz` files and dict-likes: returns the parameter tree (numpy leaves)."""
  key = None
  if isinstance(ckpt, (str, os.PathLike)):
    ckpt = str(ckpt)
    m = re.match(r"^(.*?/.*?)(?::([\w/]+))?$", ckpt)
    if m:  # '/path/to/file.npz:Encoder' -> ('/path/to/file.npz', 'Encoder')
      ckpt, key = m.groups()
    if not ckpt.endswith(".npz"):
      raise NotImplementedError("only .npz checkpoints are supported (tensorstore directories need `tensorstore`)")
    checkpoint = load_checkpoint_np(ckpt)
  else:
    checkpoint = ckpt if not _is_flat(ckpt) else load_checkpoint_np(ckpt)
  if "params" in checkpoint:
    params = checkpoint["params"]
  elif "opt" in checkpoint and isinstance(checkpoint["opt"], Mapping) and "target" in checkpoint["opt"]:
    params = checkpoint["opt"]["target"]
  else:
    params = checkpoint
  if key is not None:
    params = tree_get(params, key)
  return params


def _is_flat(d):
  return isinstance(d, Mapping) and any(isinstance(k, str) and "/" in k for k in d.keys())
