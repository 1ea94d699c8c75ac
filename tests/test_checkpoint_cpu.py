"""Checkpoint .npz import / export (big_vision/utils.py:200-290,650-706,857-884)."""
import numpy as np
import pytest
import torch

from small_vision_b200 import checkpoint as K
from small_vision_b200.config import make_model_config
from small_vision_b200.params import ArenaLayout, arena_from_tree, init_arena, tree_from_arena


def _tree():
  layout = ArenaLayout(make_model_config(variant="S/4", adaln=True, num_classes=10, channels=3, img_size=64, depth=2, dec_depth=1))
  arena = init_arena(layout, 5, "cpu", nonzero_adaln=True)
  return layout, arena, tree_from_arena(layout, arena)


def test_names_follow_sorted_depth_first_traversal():
  tree = {"b": {"y": 1, "x": 2}, "a": [3, {"k": 4}], "c": None}
  assert K.tree_flatten_with_names(tree) == [("a/0", 3), ("a/1/k", 4), ("b/x", 2), ("b/y", 1)]   # utils.py:650-673
  keys, vals = zip(*K.tree_flatten_with_names({"b": {"y": 1, "x": 2}, "d": 7}))
  assert K.recover_tree(keys, vals) == {"b": {"x": 2, "y": 1}, "d": 7}                         # utils.py:857-884


def test_param_tree_round_trips_bit_exactly(tmp_path):
  layout, arena, tree = _tree()
  path = str(tmp_path / "ckpt.npz")
  names = K.save_checkpoint_np(path, {"params": tree, "opt": {"count": np.int32(3)}})
  assert "params/Encoder/ScanCheckpointEncoder1DBlock_0/MlpBlock_0/Dense_0/kernel" in names   # Flax leaf paths (App. C)
  assert "params/final_conv/kernel" in names and "opt/count" in names
  params = K.load_params(path)                                                  # picks the "params" sub-tree
  back = arena_from_tree(layout, params, "cpu")
  assert torch.equal(back, arena)
  sub = K.load_params(path + ":Encoder/ScanCheckpointEncoder1DBlock_0")        # utils.py:258-262 sub-model suffix
  assert set(sub) == set(tree["Encoder"]["ScanCheckpointEncoder1DBlock_0"])


def test_load_params_accepts_the_three_layouts(tmp_path):
  _, _, tree = _tree()
  flat = {k: K._to_numpy(v) for k, v in K.tree_flatten_with_names(tree)}
  bare = K.load_params(dict(flat))                                              # params shared directly
  assert set(bare) == set(tree)
  old = K.load_params({"opt/target/" + k: v for k, v in flat.items()})         # Flax-optimizer checkpoints
  assert set(old) == set(tree)
  np.testing.assert_array_equal(old["final_conv"]["kernel"], bare["final_conv"]["kernel"])
  # anything that is not an .npz is a tensorstore-layout directory (utils.py:279-283; tests/test_checkpoint_ts_cpu.py)
  with pytest.raises(FileNotFoundError):
    K.load_params(str(tmp_path / "tensorstore_dir"))
  K.save_checkpoint_ts({"params": tree}, str(tmp_path / "ck"), 1)
  ts = K.load_params(str(tmp_path / "ck"))
  np.testing.assert_array_equal(ts["final_conv"]["kernel"], bare["final_conv"]["kernel"])
  with pytest.raises(KeyError):
    K.tree_get(bare, "Encoder/nope")
