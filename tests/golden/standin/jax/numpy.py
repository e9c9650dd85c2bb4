"""`jax.numpy` stand-in: numpy with float64/complex128 results narrowed to 32-bit types."""
import sys
import types

import numpy as _np

_state = {'x64': 0}


def _cast(r):
  if _state['x64']:
    return r
  if isinstance(r, _np.ndarray):
    if r.dtype == _np.float64:
      return r.astype(_np.float32)
    if r.dtype == _np.complex128:
      return r.astype(_np.complex64)
    return r
  if isinstance(r, _np.float64):
    return _np.float32(r)
  if isinstance(r, _np.complex128):
    return _np.complex64(r)
  if isinstance(r, tuple):
    return tuple(_cast(x) for x in r)
  if isinstance(r, list):
    return [_cast(x) for x in r]
  return r


def _wrap(f):
  def g(*a, **k):
    return _cast(f(*a, **k))
  g.__name__ = getattr(f, '__name__', 'fn')
  g.__wrapped__ = f
  return g


_PASS = {'finfo', 'iinfo', 'float32', 'float64', 'int32', 'int64', 'uint8', 'bool_', 'ndarray',
         'dtype', 'complex64', 'newaxis', 'pi', 'inf', 'nan', 'e', 'vectorize'}


class _Module(types.ModuleType):
  def __getattr__(self, name):
    if name.startswith('__'):
      raise AttributeError(name)
    obj = getattr(_np, name)
    if name in _PASS or not callable(obj) or isinstance(obj, type):
      return obj
    w = _wrap(obj)
    w.__name__ = name
    setattr(self, name, w)
    return w


_mod = _Module(__name__)
_mod.__dict__.update({k: v for k, v in globals().items() if k.startswith('_')})
_mod.__file__ = __file__


def _array(x, dtype=None):
  return _cast(_np.array(x, dtype=dtype))


def _copy(x):
  return _np.array(x, copy=True)


def _vectorize(f, signature=None):
  return _wrap(_np.vectorize(getattr(f, '__wrapped__', f), signature=signature))


def _matmul(a, b, precision=None):
  return _cast(_np.matmul(a, b))


_mod.matmul = _matmul
_mod.array = _array
_mod.copy = _copy
_mod.vectorize = _vectorize
_mod.linalg = types.SimpleNamespace(norm=_wrap(_np.linalg.norm), inv=_wrap(_np.linalg.inv))
sys.modules[__name__] = _mod
