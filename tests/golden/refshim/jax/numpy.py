"""numpy-fp64 stand-in for the part of jax.numpy the reference's hot path touches."""
import numpy as _np

float32 = _np.float64      # everything runs in double precision: the fixture is the exact-arithmetic answer
float64 = _np.float64
bfloat16 = _np.float64
int32 = _np.int32
dtype = _np.dtype
newaxis = None
pi = _np.pi


class _At:
  def __init__(self, arr):
    self.arr = arr

  def __getitem__(self, idx):
    arr = self.arr

    class _Set:
      def set(self, value):
        out = _np.array(arr, copy=True).view(JArr)
        out[idx] = value
        return out
    return _Set()


class JArr(_np.ndarray):
  """ndarray with the two jax.Array behaviours the reference relies on: `.at[...].set` and list-valued axes."""

  @property
  def at(self):
    return _At(self)

  def mean(self, axis=None, **kw):
    if isinstance(axis, list):
      axis = tuple(axis)
    return _np.asarray(_np.asarray(self).mean(axis=axis, **kw)).view(JArr)

  def astype(self, dt, **kw):
    return _np.asarray(self).astype(_dt(dt), **kw).view(JArr)


def _dt(d):
  if d is None:
    return None
  if isinstance(d, str):
    return {"float32": _np.float64, "bfloat16": _np.float64, "float64": _np.float64, "int32": _np.int32}[d]
  return d


def _w(a):
  return _np.asarray(a).view(JArr)


def asarray(a, dtype=None):
  a = _np.asarray(a)
  d = _dt(dtype)
  if d is None and a.dtype == _np.float32:
    d = _np.float64
  return _w(a.astype(d) if d is not None else a)


array = asarray


def zeros(shape, dtype=None):
  return _w(_np.zeros(shape, _dt(dtype) or _np.float64))


def ones(shape, dtype=None):
  return _w(_np.ones(shape, _dt(dtype) or _np.float64))


def arange(*a, step=None, dtype=None):
  if step is not None:
    return _w(_np.arange(*a, step, dtype=_dt(dtype)))
  return _w(_np.arange(*a, dtype=_dt(dtype)))


def reshape(a, shape):
  return _w(_np.reshape(a, shape))


def concatenate(xs, axis=0):
  return _w(_np.concatenate([_np.asarray(x) for x in xs], axis=axis))


def append(a, v):
  return _w(_np.append(_np.asarray(a), v))


def tile(a, reps):
  return _w(_np.tile(a, reps))


def broadcast_to(a, shape):
  return _w(_np.broadcast_to(a, shape))


def take_along_axis(a, idx, axis):
  return _w(_np.take_along_axis(_np.asarray(a), _np.asarray(idx), axis=axis))


def argsort(a, axis=-1):
  return _w(_np.argsort(_np.asarray(a), axis=axis, kind="stable"))   # jnp.argsort is stable by default


def split(a, n, axis=0):
  return [_w(p) for p in _np.split(_np.asarray(a), n, axis=axis)]


def repeat(a, n, axis=None):
  return _w(_np.repeat(_np.asarray(a), n, axis=axis))


def expand_dims(a, axis):
  return _w(_np.expand_dims(_np.asarray(a), axis))


def where(c, a, b):
  return _w(_np.where(c, a, b))


def mean(a, axis=None, keepdims=False):
  return _np.asarray(_np.asarray(a).mean(axis=axis, keepdims=keepdims)).view(JArr)


def std(a, axis=None, keepdims=False):
  return _np.asarray(_np.asarray(a).std(axis=axis, keepdims=keepdims)).view(JArr)   # population std, like jnp.std


def pad(a, pad_width, constant_values=0):
  return _w(_np.pad(_np.asarray(a), pad_width, constant_values=constant_values))


def ones_like(a):
  return _w(_np.ones_like(_np.asarray(a)))


def argmax(a, axis=None):
  return _w(_np.argmax(_np.asarray(a), axis=axis))


class linalg:  # noqa: N801
  @staticmethod
  def eigh(a):
    w, v = _np.linalg.eigh(_np.asarray(a))
    return _w(w), _w(v)


def sum(a, axis=None):  # noqa: A001
  return _np.asarray(a).sum(axis=axis)


def einsum(*a):
  return _w(_np.einsum(*a))


def sqrt(a):
  return _np.sqrt(a)


def exp(a):
  return _w(_np.exp(a))


def sin(a):
  return _w(_np.sin(a))


def cos(a):
  return _w(_np.cos(a))


def tanh(a):
  return _w(_np.tanh(a))
