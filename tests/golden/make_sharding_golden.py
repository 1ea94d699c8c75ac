"""Generates tests/golden/reference_sharding_golden.json by executing the reference's own big_vision/sharding.py
(unmodified, imported from /root/reference over tests/golden/refshim's jax.sharding / jax.tree_map value objects) on the
parameter-shape trees of the BASELINE.json model configurations.

  python tests/golden/make_sharding_golden.py       (build container only: needs /root/reference)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("UMD_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)

MODELS = {
    "umd_b4": dict(variant="B/4", adaln=True),
    "mae_b4": dict(variant="B/4", adaln=False),
    "dit_b4": dict(variant="B/4", adaln=True, num_classes=1000),
    "latent_l2": dict(variant="L/2", adaln=True, img_size=32, channels=4),
}
# (strategy, mesh size, extra args)
SETTINGS = [("replicated", 8, {}), ("fully_sharded", 8, {}), ("fully_sharded", 3, {}),
            ("fully_sharded", 8, {"too_small_to_shard_thr": 1024}), ("fully_sharded", 7, {"too_small_to_shard_thr": 0})]


class Shape:
  def __init__(self, shape):
    self.shape = tuple(shape)


def shape_tree(kw):
  """Nested dict of leaf shapes from the engine's parameter layout (Flax shapes)."""
  from small_vision_b200.model import Model
  out = {}
  for lf in Model(**kw).layout.leaves:
    d = out
    for k in lf.path[:-1]:
      d = d.setdefault(k, {})
    d[lf.path[-1]] = Shape(lf.shape)
  return out


def flatten(tree, prefix=""):
  out = {}
  for k, v in tree.items():
    if isinstance(v, dict):
      out.update(flatten(v, prefix + k + "/"))
    else:
      out[prefix + k] = v
  return out


def main():
  sys.path.insert(0, os.path.join(HERE, "refshim"))
  sys.path.insert(0, REF)
  import jax
  from big_vision import sharding as ref
  assert ref.__file__.startswith(REF) and "refshim" in jax.__file__
  gold = {"provenance": "big_vision/sharding.py executed over tests/golden/refshim", "cases": []}
  for name, kw in MODELS.items():
    tree = shape_tree(kw)
    for strategy, n, extra in SETTINGS:
      mesh = jax.sharding.Mesh(np.arange(n), ("data",))
      res = ref.infer_sharding(tree, mesh, "data", strategy, extra)
      specs = {k: list(v.spec) for k, v in flatten(res).items()}
      gold["cases"].append({"model": name, "strategy": strategy, "mesh": n, "extra": extra, "specs": specs})
    print(name, len(specs), "leaves")
  path = os.path.join(HERE, "reference_sharding_golden.json")
  json.dump(gold, open(path, "w"), indent=0, sort_keys=True)
  print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
  main()
