"""Stand-in for the part of flax.linen the reference's hot path touches (see ../README.md): a module system with
Flax's naming rules (attribute names for `setup` sub-modules, `<Class>_<n>` inside `@nn.compact`, `Scan…` /
`Checkpoint…` prefixes for lifted classes), parameters looked up in the tree given to `apply`, and the layers
restated on numpy float64.  Forward only."""
import math as _math

import numpy as _np

import jax as _jax
from jax.numpy import JArr as _JArr

_stack = []          # modules whose setup/__call__ is executing (innermost last): parents of new sub-modules
INIT_LOG = {}        # leaf path -> (module class, initialiser descriptor or None = the layer's Flax default); see Module.param
broadcast = object()
initializers = _jax.nn.initializers


def compact(fn):
  return fn


def with_logical_constraint(x, _axes):
  return x


def _w(a):
  return _np.asarray(a, dtype=_np.float64).view(_JArr)


def gelu(x, approximate=True):
  x = _np.asarray(x)
  if approximate:
    return _w(0.5 * x * (1.0 + _np.tanh(_math.sqrt(2.0 / _math.pi) * (x + 0.044715 * x ** 3))))
  from scipy.special import erf
  return _w(0.5 * x * (1.0 + erf(x / _math.sqrt(2.0))))


def silu(x):
  x = _np.asarray(x)
  return _w(x / (1.0 + _np.exp(-x)))


class Module:
  name = None

  def __init_subclass__(cls, **kw):
    super().__init_subclass__(**kw)
    call = cls.__dict__.get("__call__")
    if call is not None and not getattr(call, "_wrapped", False):
      def wrapped(self, *a, _orig=call, **k):
        self._ensure_setup()
        _stack.append(self)
        try:
          return _orig(self, *a, **k)
        finally:
          _stack.pop()
      wrapped._wrapped = True
      cls.__call__ = wrapped

  @classmethod
  def _fields(cls):
    out = {}
    for klass in reversed(cls.__mro__):
      for k in getattr(klass, "__annotations__", {}):
        if k not in ("name", "parent"):
          out[k] = None
    return list(out)

  def __init__(self, *args, **kw):
    d = object.__setattr__
    fields = self._fields()
    name = kw.pop("name", None)
    vals = dict(zip(fields, args))
    vals.update(kw)
    for k in vals:
      assert k in fields, f"{type(self).__name__}: unknown field {k}"
    for k in fields:
      if k in vals:
        d(self, k, vals[k])
      else:
        assert hasattr(type(self), k), f"{type(self).__name__}: field {k} is required"
    d(self, "_kw", vals)
    d(self, "parent", _stack[-1] if _stack else None)
    d(self, "_counters", {})
    d(self, "_in_setup", False)
    d(self, "_setup_done", False)
    d(self, "_bound", None)
    d(self, "_rngs", None)
    p = self.parent
    if name is None and p is not None and not p._in_setup:
      base = type(self).__name__
      n = p._counters.get(base, 0)
      p._counters[base] = n + 1
      name = f"{base}_{n}"
    d(self, "name", name)

  def __setattr__(self, k, v):
    if isinstance(v, Module) and self._in_setup and v.name is None:
      object.__setattr__(v, "name", k)
    object.__setattr__(self, k, v)

  def setup(self):
    pass

  def _ensure_setup(self):
    if not self._setup_done:
      object.__setattr__(self, "_setup_done", True)
      object.__setattr__(self, "_in_setup", True)
      _stack.append(self)
      try:
        self.setup()
      finally:
        _stack.pop()
        object.__setattr__(self, "_in_setup", False)

  def _params(self):
    if self._bound is not None:
      return self._bound
    assert self.parent is not None and self.name is not None, f"unbound module {type(self).__name__}"
    return self.parent._params()[self.name]

  def _path(self):
    parts, m = [], self
    while m is not None and m.parent is not None:
      parts.append(m.name)
      m = m.parent
    return tuple(reversed(parts))

  def param(self, name, _init, shape, *_a):
    path = tuple(p for p in self._path() if not p.startswith("layer")) + (name,)     # scan steps share one stacked leaf
    INIT_LOG[path] = (type(self).__name__, _init.describe() if hasattr(_init, "describe") else None)
    v = _np.asarray(self._params()[name], dtype=_np.float64)
    assert tuple(v.shape) == tuple(int(s) for s in shape), (self.name, name, v.shape, shape)
    return v.view(_JArr)

  def make_rng(self, name):
    m = self
    while m._rngs is None:
      m = m.parent
    return m._rngs[name]

  def apply(self, variables, *args, rngs=None, **kw):
    top = type(self)(**self._kw, name=self.name)
    object.__setattr__(top, "parent", None)
    object.__setattr__(top, "_bound", variables["params"])
    object.__setattr__(top, "_rngs", rngs if rngs is not None else {})
    saved = list(_stack)
    del _stack[:]
    try:
      return top(*args, **kw)
    finally:
      _stack[:] = saved


def remat(cls, **_kw):
  return type("Checkpoint" + cls.__name__, (cls,), {})


def scan(target, *, variable_axes, split_rngs, in_axes, length, **_kw):
  assert variable_axes == {"params": 0} and in_axes is broadcast

  class _Scan(Module):
    def __init__(self, **kw):
      Module.__init__(self)
      object.__setattr__(self, "_target_kw", kw)

    def __call__(self, carry, *bcast):
      stacked = self._params()

      def at(tree, i):
        return {k: (at(v, i) if isinstance(v, dict) else _np.asarray(v)[i]) for k, v in tree.items()}
      ys = []
      for i in range(length):
        blk = target(**self._target_kw, name=f"layer{i}")
        object.__setattr__(blk, "_bound", at(stacked, i))
        carry, y = blk(carry, *bcast)
        ys.append(y)
      return carry, ys

  _Scan.__name__ = _Scan.__qualname__ = "Scan" + target.__name__
  return _Scan


# ------------------------------------------------------------------------------------------------------------
# layers (library semantics restated; see README.md)
# ------------------------------------------------------------------------------------------------------------


class Dense(Module):
  features: int
  use_bias: bool = True
  dtype: object = None
  kernel_init: object = None
  bias_init: object = None

  def __call__(self, x):
    x = _np.asarray(x, dtype=_np.float64)
    k = self.param("kernel", self.kernel_init, (x.shape[-1], self.features))
    y = x @ _np.asarray(k)
    if self.use_bias:
      y = y + _np.asarray(self.param("bias", self.bias_init, (self.features,)))
    return _w(y)


class LayerNorm(Module):
  epsilon: float = 1e-6
  dtype: object = None

  def __call__(self, x):
    x = _np.asarray(x, dtype=_np.float64)
    d = x.shape[-1]
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    y = (x - mu) / _np.sqrt(var + self.epsilon)
    return _w(y * _np.asarray(self.param("scale", None, (d,))) + _np.asarray(self.param("bias", None, (d,))))


class _Proj(Module):
  """DenseGeneral as used inside MultiHeadDotProductAttention."""
  shape_in: tuple = ()
  shape_out: tuple = ()
  kernel_init: object = None

  def __call__(self, x):
    k = _np.asarray(self.param("kernel", self.kernel_init, tuple(self.shape_in) + tuple(self.shape_out)))
    b = _np.asarray(self.param("bias", None, tuple(self.shape_out)))
    ni = len(self.shape_in)
    return _np.tensordot(x, k, axes=(list(range(x.ndim - ni, x.ndim)), list(range(ni)))) + b


class MultiHeadDotProductAttention(Module):
  num_heads: int
  kernel_init: object = None
  deterministic: object = None
  dtype: object = None

  def __call__(self, inputs_q, inputs_kv):
    xq = _np.asarray(inputs_q, dtype=_np.float64)
    xkv = _np.asarray(inputs_kv, dtype=_np.float64)
    d = xq.shape[-1]
    h = self.num_heads
    assert d % h == 0
    hd = d // h
    q = _Proj((d,), (h, hd), self.kernel_init, name="query")(xq)      # [b, s, h, hd]
    k = _Proj((d,), (h, hd), self.kernel_init, name="key")(xkv)
    v = _Proj((d,), (h, hd), self.kernel_init, name="value")(xkv)
    out = _np.empty_like(q)
    for b in range(q.shape[0]):
      for j in range(h):
        logits = (q[b, :, j] / _math.sqrt(hd)) @ k[b, :, j].T
        logits = logits - logits.max(-1, keepdims=True)
        w = _np.exp(logits)
        w = w / w.sum(-1, keepdims=True)
        out[b, :, j] = w @ v[b, :, j]
    return _w(_Proj((h, hd), (d,), self.kernel_init, name="out")(out))


class Dropout(Module):
  rate: float = 0.0

  def __call__(self, x, deterministic=None):
    assert self.rate == 0.0 or deterministic, "dropout > 0 is not used by any recipe of the hot path"
    return x


class Conv(Module):
  features: int
  kernel_size: tuple = ()
  strides: tuple = None
  padding: str = "SAME"
  dtype: object = None
  kernel_init: object = None

  def __call__(self, x):
    x = _np.asarray(x, dtype=_np.float64)
    kh, kw = self.kernel_size
    assert tuple(self.strides) == (kh, kw) and self.padding == "VALID"
    n, H, W, c = x.shape
    k = _np.asarray(self.param("kernel", self.kernel_init, (kh, kw, c, self.features)))
    b = _np.asarray(self.param("bias", None, (self.features,)))
    out = _np.zeros((n, H // kh, W // kw, self.features))
    for a in range(kh):                 # cross-correlation: out[i, j] = sum_ab x[i*kh + a, j*kw + b] K[a, b]
      for bb in range(kw):
        out += x[:, a::kh, bb::kw, :][:, :H // kh, :W // kw] @ k[a, bb]
    return _w(out + b)


class ConvTranspose(Module):
  features: int
  kernel_size: tuple = ()
  strides: tuple = None
  padding: str = "SAME"
  dtype: object = None
  kernel_init: object = None

  def __call__(self, x):
    """jax.lax.conv_transpose, transpose_kernel=False: dilate the input by the stride, pad by k-1 on both sides
    ('VALID'), then cross-correlate with the kernel as stored (no flip, no in/out swap)."""
    x = _np.asarray(x, dtype=_np.float64)
    kh, kw = self.kernel_size
    sh, sw = self.strides
    assert self.padding == "VALID"
    n, h, w, c = x.shape
    k = _np.asarray(self.param("kernel", self.kernel_init, (kh, kw, c, self.features)))
    b = _np.asarray(self.param("bias", None, (self.features,)))
    Hd, Wd = (h - 1) * sh + 1, (w - 1) * sw + 1
    xd = _np.zeros((n, Hd + 2 * (kh - 1), Wd + 2 * (kw - 1), c))
    xd[:, kh - 1:kh - 1 + Hd:sh, kw - 1:kw - 1 + Wd:sw, :] = x
    Ho, Wo = Hd + kh - 1, Wd + kw - 1
    # 'VALID' for a transposed convolution also pads max(stride - kernel, 0) at the end; zero here (stride = kernel)
    assert sh <= kh and sw <= kw
    out = _np.zeros((n, Ho, Wo, self.features))
    for a in range(kh):
      for bb in range(kw):
        out += xd[:, a:a + Ho, bb:bb + Wo, :] @ k[a, bb]
    return _w(out + b)


class Embed(Module):
  num_embeddings: int
  features: int = 0

  def __call__(self, ids):
    table = _np.asarray(self.param("embedding", None, (self.num_embeddings, self.features)))
    return _w(table[_np.asarray(ids).astype(_np.int64)])
