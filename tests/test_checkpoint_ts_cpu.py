"""The tensorstore-layout checkpoint directories of big_vision/utils.py:886-1016: directory and
array naming ('/' -> '~'), the `-LAST` pointer / `-tmp` protocol, the zarr-v2 array stores, and loading through
`load_params`.  Where /root/reference is present (the build container) the reference's own `save_checkpoint_ts` /
`load_checkpoint_ts` / `tssave` / `tsload` are lifted with `ast` and run over a stand-in for the tensorstore manager that
records the paths they ask for: same directories, same pointer files, same names as ours.  The bytes of the array stores
follow the zarr v2 specification (tensorstore itself is not installable here: parity unpinned for the bytes)."""
import ast
import functools
import json
import os
import re
import shutil
import sys

import numpy as np
import pytest

from small_vision_b200 import checkpoint as CK

REF = os.environ.get("UMD_REFERENCE_ROOT", "/root/reference")


def _tree(seed=0):
  g = np.random.default_rng(seed)
  return {"params": {"Encoder": {"blk": {"kernel": g.standard_normal((3, 8, 5)).astype(np.float32),
                                         "bias": g.standard_normal((3, 5)).astype(np.float32)},
                                 "encoder_norm": {"scale": np.ones(8, np.float32)}},
                     "cls": g.standard_normal((1, 4, 8)).astype(np.float32)},
          "opt": {"count": np.asarray(7, np.int32), "nu": {"cls": g.standard_normal((1, 4, 8)).astype(np.float32)}}}


def _flat(t, p=""):
  out = {}
  for k, v in t.items():
    if isinstance(v, dict):
      out.update(_flat(v, p + k + "/"))
    else:
      out[p + k] = v
  return out


@pytest.mark.parametrize("compressor", [None, {"id": "zstd", "level": 1}])
def test_ts_checkpoint_round_trip_and_pointer_protocol(tmp_path, compressor):
  base = str(tmp_path / "checkpoint.bv")
  t0, t1, t2 = _tree(0), _tree(1), _tree(2)
  assert CK.save_checkpoint_ts(t0, base, 10, keep=False, compressor=compressor) == "000000010-tmp"
  assert open(base + "-LAST").read() == "000000010-tmp"
  d = base + "-000000010-tmp"
  assert sorted(os.listdir(d)) == sorted(n.replace("/", "~") for n in _flat(t0))
  assert "params~Encoder~blk~kernel" in os.listdir(d)
  meta = json.load(open(os.path.join(d, "params~Encoder~blk~kernel", ".zarray")))
  assert meta["shape"] == [3, 8, 5] and meta["chunks"] == [3, 8, 5] and meta["dtype"] == "<f4" and meta["zarr_format"] == 2
  assert os.path.exists(os.path.join(d, "params~Encoder~blk~kernel", "0.0.0"))
  assert os.path.exists(os.path.join(d, "opt~count", "0"))                     # rank-0 array
  got = CK.load_checkpoint_ts(base)
  for k, v in _flat(t0).items():
    assert np.array_equal(_flat(got)[k], v) and _flat(got)[k].dtype == v.dtype
  # the next save removes the temporary predecessor, a kept one stays
  assert CK.save_checkpoint_ts(t1, base, 20, keep=True, compressor=compressor) == "000000020"
  assert not os.path.exists(d) and open(base + "-LAST").read() == "000000020"
  CK.save_checkpoint_ts(t2, base, 30, keep=False, compressor=compressor)
  assert os.path.isdir(base + "-000000020") and os.path.isdir(base + "-000000030-tmp")
  got = CK.load_checkpoint_ts(base)
  assert np.array_equal(got["params"]["cls"], t2["params"]["cls"])
  # a specific step directory, a regex filter, an explicit tree
  got = CK.load_checkpoint_ts(base + "-000000020", regex="params/.*")
  assert set(got) == {"params"} and np.array_equal(got["params"]["cls"], t1["params"]["cls"])
  got = CK.tsload(base + "-000000020", tree={"opt": {"count": 0}})
  assert int(got["opt"]["count"]) == 7
  with pytest.raises(ValueError):
    CK.tsload(base + "-000000020", tree={"a": 0}, regex="x")
  with pytest.raises(ValueError, match="~"):
    CK.tssave({"a~b": np.zeros(2)}, str(tmp_path / "bad"))
  # load_params: whole parameter tree, and the ':sub/key' suffix (utils.py:262,279-283)
  p = CK.load_params(base)
  assert np.array_equal(p["Encoder"]["blk"]["kernel"], t2["params"]["Encoder"]["blk"]["kernel"])
  sub = CK.load_params(base + ":Encoder/blk")
  assert sorted(sub) == ["bias", "kernel"]


def test_zarr_reader_handles_chunk_grids_and_separators(tmp_path):
  """Arrays written by a sharded run have several chunks (one per shard); '/' separated keys are legal zarr v2."""
  a = np.arange(7 * 5, dtype=np.float32).reshape(7, 5)
  for sep in (".", "/"):
    d = str(tmp_path / f"arr{ord(sep)}")
    os.makedirs(d)
    json.dump({"chunks": [4, 2], "compressor": None, "dtype": "<f4", "fill_value": 0.0, "filters": None, "order": "C",
               "shape": [7, 5], "zarr_format": 2, "dimension_separator": sep}, open(os.path.join(d, ".zarray"), "w"))
    for i in range(2):
      for j in range(3):
        if (i, j) == (1, 2):
          continue                                  # a missing chunk reads as fill_value
        blk = np.zeros((4, 2), np.float32)
        sub = a[i * 4:(i + 1) * 4, j * 2:(j + 1) * 2]
        blk[:sub.shape[0], :sub.shape[1]] = sub
        fn = os.path.join(d, *f"{i}{sep}{j}".split("/"))
        os.makedirs(os.path.dirname(fn), exist_ok=True)
        open(fn, "wb").write(blk.tobytes())
    got = CK._zarr_read(d)
    want = a.copy()
    want[4:, 4:] = 0
    assert np.array_equal(got, want)


def test_arena_loads_from_a_ts_checkpoint(tmp_path):
  """A parameter tree written from the (strided, layer-major) arena views comes back bit for bit."""
  import torch
  from small_vision_b200.model import Model
  from small_vision_b200.params import arena_from_tree, init_arena, tree_from_arena
  model = Model(variant="S/4", adaln=True, num_classes=10, depth=2, dec_depth=1)
  arena = init_arena(model.layout, 3, "cpu", nonzero_adaln=True)
  tree = tree_from_arena(model.layout, arena)
  base = str(tmp_path / "ck")
  CK.save_checkpoint_ts({"params": tree, "opt": {"count": np.asarray(3)}}, base, 5)
  back = arena_from_tree(model.layout, CK.load_params(base), "cpu")
  assert torch.equal(back, arena)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference sources only exist in the build container")
def test_directory_protocol_matches_the_references_own_functions(tmp_path):
  """big_vision/utils.py's tssave / save_checkpoint_ts / load_checkpoint_ts / tsload (lifted, unmodified) over a
  stand-in for jax's GlobalAsyncCheckpointManager and array_serialization that stores with OUR zarr writer / reader:
  the reference produces the same directory names, array names and pointer files as save_checkpoint_ts here, and its
  loader reads back what ours wrote."""
  import collections
  import multiprocessing.pool
  import types
  here = os.path.dirname(os.path.abspath(__file__))
  sys.path.insert(0, os.path.join(here, "golden", "refshim"))
  import jax
  assert "refshim" in jax.__file__
  path = os.path.join(REF, "big_vision", "utils.py")
  src = ast.parse(open(path).read(), filename=path)
  want = ("tssave", "save_checkpoint_ts", "load_checkpoint_ts", "tsload", "tree_flatten_with_names", "_traverse_with_names",
          "recover_tree", "tree_broadcast")

  class _Gfile:
    makedirs = staticmethod(lambda p: os.makedirs(p, exist_ok=True))
    exists = staticmethod(os.path.exists)
    listdir = staticmethod(os.listdir)
    rmtree = staticmethod(lambda p: shutil.rmtree(p, ignore_errors=True))
    GFile = staticmethod(lambda p, mode="r": open(p, mode))

    @staticmethod
    def rename(a, b, overwrite=False):
      os.replace(a, b)

  class _Mngr:
    def serialize_with_paths(self, vals, paths, on_commit_callback):
      for v, p in zip(vals, paths):
        CK._zarr_write(p, np.asarray(v))
      on_commit_callback()

  class _Ser:
    @staticmethod
    def get_tensorstore_spec(p):
      return p

    @staticmethod
    def run_deserialization(shardings, specs):
      return [CK._zarr_read(p) for p in specs]

  class _Guard:
    def __enter__(self): return self
    def __exit__(self, *a): return False

  jx = types.SimpleNamespace(transfer_guard=lambda *_: _Guard(), tree_leaves=lambda t: list(_flat(t).values()) if isinstance(t, dict) else [t],
                             tree_util=jax.tree_util, devices=lambda *_: [0],
                             sharding=types.SimpleNamespace(SingleDeviceSharding=lambda d: 0))
  env = dict(os=os, re=re, np=np, functools=functools, collections=collections, multiprocessing=multiprocessing, jax=jx, gfile=_Gfile,
             array_serial=_Ser, flax=types.SimpleNamespace(), mlc=types.SimpleNamespace(), dataclasses=__import__("dataclasses"),
             Mapping=__import__("collections.abc").abc.Mapping)
  for node in src.body:
    if isinstance(node, ast.FunctionDef) and node.name in want:
      exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), env)
  env["tree_broadcast"] = lambda prefix, tree: {k: prefix for k in _flat(tree)}    # shardings are irrelevant on one host
  ref_base, our_base = str(tmp_path / "ref" / "ck"), str(tmp_path / "ours" / "ck")
  os.makedirs(os.path.dirname(ref_base)); os.makedirs(os.path.dirname(our_base))
  for step, keep, seed in ((10, False, 0), (20, True, 1), (30, False, 2)):
    env["save_checkpoint_ts"](_Mngr(), _tree(seed), ref_base, step, keep=keep)
    CK.save_checkpoint_ts(_tree(seed), our_base, step, keep=keep)

  def listing(base):
    root = os.path.dirname(base)
    return sorted(os.path.relpath(os.path.join(d, f), root) for d, _, fs in os.walk(root) for f in fs)

  assert listing(ref_base) == listing(our_base)
  assert open(ref_base + "-LAST").read() == open(our_base + "-LAST").read() == "000000030-tmp"
  got = env["load_checkpoint_ts"](our_base)                       # the reference's loader on OUR directory
  for k, v in _flat(_tree(2)).items():
    assert np.array_equal(_flat(got)[k], v)
  ours = CK.load_checkpoint_ts(ref_base)                          # our loader on the reference-written directory
  for k, v in _flat(_tree(2)).items():
    assert np.array_equal(_flat(ours)[k], v)
