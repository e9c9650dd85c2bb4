"""numpy stand-in for the parts of `jax` the reference's per-ray math uses (see ../README.md)."""
import numpy as _np

from . import numpy  # noqa: F401  (jax.numpy)
from . import nn, lax, random, tree_util  # noqa: F401
from .numpy import _cast, _state

_reentry_hooks = []   # objects with save()/restore(state): see flax stand-in


def _save_hooks():
  return [(h, h.save()) for h in _reentry_hooks]


def _restore_hooks(saved):
  for h, st in saved:
    h.restore(st)


def vmap(fn, in_axes=0, out_axes=0):
  def mapped(*args):
    axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
    n = None
    moved = []
    for a, ax in zip(args, axes):
      if ax is None:
        moved.append(None)
      else:
        a = _np.asarray(a)
        moved.append(_np.moveaxis(a, ax, 0))
        n = a.shape[ax]
    outs = []
    saved = _save_hooks()
    for i in range(n):
      call = [a if m is None else m[i] for a, m in zip(args, moved)]
      if i > 0:
        _restore_hooks(saved)
      outs.append(fn(*call))
    if isinstance(outs[0], tuple):
      return tuple(_stack([o[j] for o in outs], out_axes) for j in range(len(outs[0])))
    return _stack(outs, out_axes)
  return mapped


def _stack(items, axis):
  if isinstance(items[0], tuple):
    return tuple(_stack([it[j] for it in items], axis) for j in range(len(items[0])))
  arr = _np.stack([_np.asarray(x) for x in items], axis=0)
  return _np.moveaxis(arr, 0, axis)


class custom_jvp:
  def __init__(self, fn):
    self.fn = fn
    self.__name__ = getattr(fn, '__name__', 'custom_jvp')

  def defjvp(self, jvp):
    self.jvp = jvp
    return jvp

  def __call__(self, *a, **k):
    return self.fn(*a, **k)


def linearize(fn, primal):
  """(fn(primal), v -> J v) with J v by fp64 central differences along the unit direction of v
  (per last-axis vector), step 1e-6 * max(1, |primal|): covariance columns can be 1e11 long while
  the primal is O(1e6), so the step must not scale with |v|."""
  out = fn(primal)
  p64 = _np.asarray(primal, dtype=_np.float64)

  def lin(v):
    v64 = _np.asarray(v, dtype=_np.float64)
    nv = _np.linalg.norm(v64, axis=-1, keepdims=True)
    vhat = v64 / _np.maximum(nv, 1e-300)
    h = 1e-6 * _np.maximum(1.0, _np.linalg.norm(p64, axis=-1, keepdims=True))
    _state['x64'] += 1
    try:
      d = (_np.asarray(fn(p64 + h * vhat), dtype=_np.float64) -
           _np.asarray(fn(p64 - h * vhat), dtype=_np.float64)) / (2 * h)
    finally:
      _state['x64'] -= 1
    return _cast(d * nv)
  return out, lin


def value_and_grad(fn, argnums=0, has_aux=False):
  """Scalar-output fn; gradient w.r.t. args[argnums] by fp64 central differences."""
  def vg(*args):
    saved = _save_hooks()
    val = fn(*args)
    after = _save_hooks()
    x64 = _np.asarray(args[argnums], dtype=_np.float64)
    g = _np.zeros_like(x64)
    h = 1e-7        # fp64 probes: IPE scales reach 2^15, the step must stay far below 2^-15
    _state['x64'] += 1
    try:
      for i in range(x64.size):
        e = _np.zeros_like(x64).reshape(-1)
        e[i] = h
        e = e.reshape(x64.shape)
        ap = list(args); am = list(args)
        ap[argnums] = x64 + e; am[argnums] = x64 - e
        _restore_hooks(saved)
        fp = fn(*ap)
        _restore_hooks(saved)
        fm = fn(*am)
        if has_aux:
          fp, fm = fp[0], fm[0]
        g.reshape(-1)[i] = (float(fp) - float(fm)) / (2 * h)
    finally:
      _state['x64'] -= 1
      _restore_hooks(after)
    return val, _cast(g)
  return vg


def host_id():
  return 0


def process_index():
  return 0


def process_count():
  return 1


def device_count():
  return 1


def local_device_count():
  return 1


def jit(fn, *a, **k):
  return fn


def pmap(fn, *a, **k):
  return fn
