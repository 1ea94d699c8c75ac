"""Generates tests/golden/reference_checkpoint_golden.json by running the reference's own checkpoint readers on files
written by small-vision_b200/checkpoint.py.  `npload`, `load_checkpoint_np`, `load_params`, `recover_tree`, `tree_get`,
`_traverse_with_names` and `tree_flatten_with_names` are lifted out of /root/reference/big_vision/utils.py with `ast`
(utils.py as a whole imports tensorflow / ml_collections and cannot be imported here); they only need os / io / re /
numpy / collections plus `jax.tree_util.tree_flatten`, which tests/golden/refshim restates for dicts.  At generation
time the script ASSERTS that the reference readers return exactly the tree that was written, for the three container
shapes `load_params` recognises ("params", "opt/target", bare tree) and for the `file.npz:sub/key` suffix; the JSON
records the reference's leaf names (order included) and per-leaf digests for the CPU test.

  python tests/golden/make_checkpoint_golden.py       (build container only: needs /root/reference)
"""
import ast
import collections
import dataclasses
import hashlib
import io
import json
import os
import re
import sys
import tempfile
import types
from typing import Mapping

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("UMD_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)

MODEL = dict(variant="S/4", adaln=True, num_classes=10, depth=2, dec_depth=1)
LIFT = ("npload", "load_checkpoint_np", "load_params", "recover_tree", "tree_get", "_traverse_with_names",
        "tree_flatten_with_names")


def make_tree(seed=0):
  """Seeded numpy parameter tree with the model's Flax leaf paths and shapes (float32)."""
  from small_vision_b200.model import Model
  rng = np.random.default_rng(seed)
  tree = {}
  for lf in Model(**MODEL).layout.init_order:   # the seeded order the fixture was generated with
    d = tree
    for k in lf.path[:-1]:
      d = d.setdefault(k, {})
    d[lf.path[-1]] = rng.standard_normal(lf.shape).astype(np.float32)
  return tree


def containers(tree, seed=1):
  rng = np.random.default_rng(seed)
  opt_like = {"count": np.asarray(7, np.int32), "mu": {"cls": rng.standard_normal((1, 4, 384)).astype(np.float32)}}
  return {"bare": tree, "params": {"params": tree, "opt": opt_like}, "flax_opt": {"opt": {"target": tree, "state": opt_like}}}


def digest(a):
  a = np.ascontiguousarray(a)
  return hashlib.sha256(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes()).hexdigest()[:16]


def lift_reference():
  sys.path.insert(0, os.path.join(HERE, "refshim"))
  import jax
  assert "refshim" in jax.__file__
  path = os.path.join(REF, "big_vision", "utils.py")
  tree = ast.parse(open(path).read(), filename=path)
  env = dict(os=os, io=io, re=re, np=np, collections=collections, dataclasses=dataclasses, Mapping=Mapping, jax=jax,
             gfile=types.SimpleNamespace(), flax=types.SimpleNamespace(), mlc=types.SimpleNamespace())
  lines = {}
  for name in LIFT:
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), env)
    lines[name] = (node.lineno, node.end_lineno)
  return env, lines


def flat(tree, prefix=""):
  out = {}
  for k, v in tree.items():
    if isinstance(v, dict):
      out.update(flat(v, prefix + k + "/"))
    else:
      out[prefix + k] = v
  return out


def main():
  from small_vision_b200 import checkpoint as CK
  ref, lines = lift_reference()
  tree = make_tree()
  want = flat(tree)
  gold = {"provenance": "readers of big_vision/utils.py (lines %s) run on files written by small-vision_b200/checkpoint.py" % lines,
          "names": [n for n, _ in ref["tree_flatten_with_names"](tree)[0]],
          "digests": {k: digest(v) for k, v in want.items()}, "containers": {}}
  assert gold["names"] == [n for n, _ in CK.tree_flatten_with_names(tree)], "leaf naming / order differs from the reference"
  with tempfile.TemporaryDirectory() as tmp:
    for cname, obj in containers(tree).items():
      path = os.path.join(tmp, cname + ".npz")
      written = CK.save_checkpoint_np(path, obj)
      got = ref["load_params"](path)                                     # the reference reads our file
      gf = flat(got)
      assert set(gf) == set(want), sorted(set(gf) ^ set(want))[:5]
      assert all(gf[k].dtype == want[k].dtype and np.array_equal(gf[k], want[k]) for k in want)
      sub = ref["load_params"](path + ":Encoder/encoder_norm")           # the ':sub/key' suffix (utils.py:262)
      assert sorted(sub) == ["bias", "scale"] and np.array_equal(sub["scale"], tree["Encoder"]["encoder_norm"]["scale"])
      leaf = ref["load_params"](path + ":cls")
      assert np.array_equal(leaf, tree["cls"])
      full = ref["load_checkpoint_np"](path)                             # whole container, optimiser entries included
      gold["containers"][cname] = {"file_keys": written, "top_level": sorted(full),
                                   "all_digests": {k: digest(v) for k, v in flat(full).items()}}
      print(cname, len(written), "arrays; reference readers round-trip OK")
  out = os.path.join(HERE, "reference_checkpoint_golden.json")
  json.dump(gold, open(out, "w"), indent=0, sort_keys=True)
  print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
  main()
