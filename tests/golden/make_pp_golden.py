"""Generates tests/golden/reference_pp_golden.npz by executing the reference's own `get_value_range`
(big_vision/pp/ops_general.py:30-62, lifted with `ast`, decorators dropped — they only route dict keys) over a
numpy-float32 stand-in for the five TensorFlow names it uses (tf.constant, tf.cast, tf.float32, tf.clip_by_value and
tensor arithmetic).  This pins the operation order of the value-range stage; the bilinear resize itself is TensorFlow's
arithmetic (absent) and stays a restatement (oracle/umd_oracle.py::preprocess_train).

  python tests/golden/make_pp_golden.py       (build container only: needs /root/reference)
"""
import ast
import os
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("UMD_REFERENCE_ROOT", "/root/reference")

# (vmin, vmax, in_min, in_max, clip_values)
SETTINGS = [(-1, 1, 0, 255.0, False), (0, 1, 0, 255.0, False), (-1, 1, 16, 235.0, True), (-2.5, 0.5, 0, 255.0, False), (0.1, 0.7, 0, 255.0, False)]


def all_bytes():
  """Every uint8 value in every channel position: [1, 16, 16, 3]."""
  return np.stack([np.arange(256, dtype=np.uint8).reshape(16, 16)] * 3, axis=-1)[None]


def tf_standin():
  f32 = np.float32
  return types.SimpleNamespace(
      float32=f32,
      constant=lambda v, dtype=None: np.asarray(v, dtype=dtype),
      cast=lambda x, dtype: np.asarray(x).astype(dtype),
      clip_by_value=lambda x, lo, hi: np.clip(x, f32(lo), f32(hi)).astype(f32))


def main():
  path = os.path.join(REF, "big_vision", "pp", "ops_general.py")
  tree = ast.parse(open(path).read(), filename=path)
  node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_value_range")
  node.decorator_list = []
  env = {"tf": tf_standin()}
  exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), env)
  x = all_bytes()
  out = {}
  for i, (vmin, vmax, in_min, in_max, clip) in enumerate(SETTINGS):
    y = env["get_value_range"](vmin, vmax, in_min, in_max, clip)(x)
    # python scalars times float32 arrays stay float32 in numpy, as they do for tf tensors
    assert y.dtype == np.float32, y.dtype
    out[f"case{i}"] = y
  dst = os.path.join(HERE, "reference_pp_golden.npz")
  np.savez_compressed(dst, **out)
  print(dst, os.path.getsize(dst), "bytes; lines", node.lineno, node.end_lineno)


if __name__ == "__main__":
  main()
