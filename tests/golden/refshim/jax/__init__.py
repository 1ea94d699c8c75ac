"""Stand-in for the `jax` top-level names the reference's hot path touches (see ../README.md)."""
import types as _types

import numpy as _np

from . import numpy  # noqa: F401  (jax.numpy)
from . import sharding  # noqa: F401  (jax.sharding)
from .numpy import JArr as _JArr


class Key:
  """A PRNG key that carries the draws it will hand out: `draws` maps a kind ('uniform', 'normal',
  'bernoulli') to an array, or to a list of arrays consumed in call order.  `split` hands the same
  dictionary on under a derived name so that supplied draws can be addressed per consumer."""

  def __init__(self, draws=None, name="root"):
    self.draws = draws if draws is not None else {}
    self.name = name

  def take(self, kind, shape):
    v = self.draws[kind]
    if isinstance(v, list):
      v = v.pop(0)
    v = _np.asarray(v)
    assert tuple(v.shape) == tuple(shape), (kind, v.shape, shape)
    return v.view(_JArr)


def _split(key, n=2):
  subs = key.draws.get("split")
  if subs is not None:
    assert len(subs) == n
    return list(subs)
  return [key for _ in range(n)]   # list-valued draws are then consumed in call order


random = _types.SimpleNamespace(
    split=_split,
    uniform=lambda key, shape: key.take("uniform", shape),
    normal=lambda key, shape: key.take("normal", shape),
    bernoulli=lambda key, p, shape: key.take("bernoulli", shape).astype(bool),
    randint=lambda key, shape, dtype=None, minval=0, maxval=None: key.take("randint", shape),
)


def vmap(fn):
  def run(*args):
    return _np.stack([_np.asarray(fn(*[a[i] for a in args])) for i in range(len(args[0]))]).view(_JArr)
  return run


def _scan(f, init, xs):
  carry = init
  for x in xs:
    carry, _ = f(carry, x)
  return carry, None


lax = _types.SimpleNamespace(scan=_scan)
checkpoint_policies = _types.SimpleNamespace(nothing_saveable=None)


class InitDesc:
  """An initialiser as the reference names it: `nn.initializers.zeros` (uncalled) or `nn.initializers.normal(stddev=..)`
  (called).  Never evaluated — parameters are supplied to `apply` — but recorded per leaf by flax.linen.Module.param."""

  def __init__(self, name, args=(), kwargs=None):
    self.name, self.args, self.kwargs = name, tuple(args), dict(kwargs or {})

  def __call__(self, *args, **kwargs):
    return InitDesc(self.name, args, kwargs)

  def describe(self):
    num = lambda v: isinstance(v, (int, float)) or (hasattr(v, "dtype") and getattr(v, "ndim", 1) == 0)
    return {"name": self.name, "args": [float(a) for a in self.args if num(a)],
            "kwargs": {k: float(v) for k, v in self.kwargs.items() if num(v)}}


class _Init:
  def __getattr__(self, name):
    return InitDesc(name)


def _one_hot(y, num_classes):
  y = _np.asarray(y).astype(_np.int64)
  return (y[..., None] == _np.arange(num_classes)).astype(_np.float64).view(_JArr)


nn = _types.SimpleNamespace(initializers=_Init(), one_hot=_one_hot)


def tree_map(fn, tree):
  """jax.tree_map over nested dicts (the only containers in the parameter tree)."""
  if isinstance(tree, dict):
    return {k: tree_map(fn, v) for k, v in tree.items()}
  return fn(tree)


class _TreeDef:
  def __init__(self, skeleton):
    self.skeleton = skeleton

  def unflatten(self, leaves):
    it = iter(leaves)

    def build(node):
      if isinstance(node, dict):
        return {k: build(node[k]) for k in sorted(node)}
      if isinstance(node, (list, tuple)):
        return type(node)(build(v) for v in node)
      return next(it)
    return build(self.skeleton)


def _tree_flatten(tree):
  """jax.tree_util.tree_flatten for dicts (keys in sorted order), lists and tuples; None is an empty sub-tree."""
  leaves = []

  def walk(node):
    if node is None:
      return None
    if isinstance(node, dict):
      return {k: walk(node[k]) for k in sorted(node)}
    if isinstance(node, (list, tuple)):
      return type(node)(walk(v) for v in node)
    leaves.append(node)
    return 0
  return leaves, _TreeDef(walk(tree))


tree_util = _types.SimpleNamespace(tree_flatten=_tree_flatten, tree_map=tree_map, tree_leaves=lambda t: _tree_flatten(t)[0])


def jit(fn=None, **_kw):
  """jax.jit as a pass-through (also when used through functools.partial(jax.jit, static_argnums=...))."""
  return fn if fn is not None else (lambda f: f)

