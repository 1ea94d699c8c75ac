"""small-vision_b200/checkpoint.py against tests/golden/reference_checkpoint_golden.json: files written by our writer were
read back by the reference's own `load_params` / `load_checkpoint_np` / `tree_flatten_with_names` (lifted from
big_vision/utils.py, tests/golden/make_checkpoint_golden.py — the round trip is asserted at generation time); here our
readers must report the same leaf names, order, container detection and per-leaf bytes as the reference's did."""
import json
import os

import numpy as np

from small_vision_b200 import checkpoint as CK
from tests.golden import make_checkpoint_golden as CG

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_checkpoint_golden.json")))


def test_leaf_names_and_order_match_the_reference_flattening():
  tree = CG.make_tree()
  assert [n for n, _ in CK.tree_flatten_with_names(tree)] == GOLD["names"]
  assert {k: CG.digest(v) for k, v in CG.flat(tree).items()} == GOLD["digests"]      # the fixture's tree regenerates


def test_our_readers_agree_with_the_reference_readers(tmp_path):
  tree = CG.make_tree()
  for cname, obj in CG.containers(tree).items():
    want = GOLD["containers"][cname]
    path = str(tmp_path / (cname + ".npz"))
    assert CK.save_checkpoint_np(path, obj) == want["file_keys"]
    assert sorted(np.load(path).files) == sorted(want["file_keys"])
    full = CK.load_checkpoint_np(path)
    assert sorted(full) == want["top_level"]
    assert {k: CG.digest(v) for k, v in CG.flat(full).items()} == want["all_digests"]
    params = CK.load_params(path)                                   # "params" / "opt/target" / bare detection
    assert {k: CG.digest(v) for k, v in CG.flat(params).items()} == GOLD["digests"]
    sub = CK.load_params(path + ":Encoder/encoder_norm")
    assert sorted(sub) == ["bias", "scale"]
    assert np.array_equal(CK.load_params(path + ":cls"), tree["cls"])
