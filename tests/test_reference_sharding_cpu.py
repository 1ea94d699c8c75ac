"""small-vision_b200/sharding.py::infer_sharding against tests/golden/reference_sharding_golden.json — the decisions of
the reference's own big_vision/sharding.py (executed over the jax stand-in, tests/golden/make_sharding_golden.py) on the
parameter-shape trees of the BASELINE.json model configurations, both strategies, divisible and non-divisible meshes."""
import json
import os

import numpy as np
import pytest

from small_vision_b200.sharding import Mesh, infer_sharding
from tests.golden import make_sharding_golden as SG

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_sharding_golden.json")))


@pytest.mark.parametrize("model", sorted(SG.MODELS))
def test_infer_sharding_matches_reference_source(model):
  tree = SG.shape_tree(SG.MODELS[model])
  cases = [c for c in GOLD["cases"] if c["model"] == model]
  assert len(cases) == len(SG.SETTINGS)
  n_sharded = 0
  for c in cases:
    mesh = Mesh(np.arange(c["mesh"]), ("data",))
    got = SG.flatten(infer_sharding(tree, mesh, "data", c["strategy"], c["extra"]))
    assert set(got) == set(c["specs"])
    for k, want in c["specs"].items():
      assert list(got[k]) == want, (model, c["strategy"], c["mesh"], k, got[k], want)
      n_sharded += any(p is not None for p in want)
  assert n_sharded > 0
