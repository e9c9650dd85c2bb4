"""Minimal flax.linen: Module with dataclass-style fields, @compact auto-naming (ClassName_i per
parent, in creation order), Dense / Embed reading parameters from the tree given to apply()."""
import numpy as _np

import jax
from jax import nn as _jnn

relu = _jnn.relu
softplus = _jnn.softplus
sigmoid = _jnn.sigmoid
silu = _jnn.silu

_STACK = []          # modules whose __call__ is running
_GIN = {}            # class name -> {field: value}; emulates gin bindings (see gin stand-in)


class _Counters:
  """Re-entrancy hook for the jax stand-in: vmap / value_and_grad call a function many times,
  flax (tracing once) names submodules once -- restore the name counters between calls."""

  @staticmethod
  def save():
    return [(m, dict(m._counters)) for m in _STACK]

  @staticmethod
  def restore(state):
    for m, c in state:
      m._counters = dict(c)


jax._reentry_hooks.append(_Counters)


def compact(fn):
  fn._compact = True
  return fn


class Module:
  def __init_subclass__(cls, **kw):
    super().__init_subclass__(**kw)

  def __init__(self, *args, name=None, parent=None, **kwargs):
    fields = {}
    for klass in reversed(type(self).__mro__):
      for k in getattr(klass, '__annotations__', {}):
        if hasattr(klass, k):
          fields[k] = getattr(klass, k)
        elif k not in fields:
          fields[k] = None
    names = [k for k in fields]
    for k, v in zip(names, args):
      fields[k] = v
    for klass in type(self).__mro__:
      for k, v in _GIN.get(klass.__name__, {}).items():
        if k in fields and klass.__name__ == type(self).__name__:
          fields[k] = v
    fields.update(kwargs)
    for k, v in fields.items():
      object.__setattr__(self, k, v)
    self._counters = {}
    self._setup_done = False
    self._parent = _STACK[-1] if _STACK else None
    if self._parent is not None:
      if name is None:
        base = type(self).__name__
        i = self._parent._counters.get(base, 0)
        self._parent._counters[base] = i + 1
        name = f'{base}_{i}'
      self._name = name
      self._params = self._parent._params[name]
    else:
      self._name = name
      self._params = None

  def setup(self):
    pass

  def _run(self, *args, **kwargs):
    if not self._setup_done:
      self._setup_done = True
      self.setup()
    _STACK.append(self)
    saved = dict(self._counters)
    try:
      return type(self).__call__(self, *args, **kwargs)
    finally:
      _STACK.pop()
      if self._parent is not None or True:
        pass

  def apply(self, variables, *args, **kwargs):
    self._params = variables['params'] if 'params' in variables else variables
    self._counters = {}
    return self._run(*args, **kwargs)

  def __getattribute__(self, item):
    return object.__getattribute__(self, item)


def _wrap_call(cls):
  orig = cls.__dict__.get('__call__')
  if orig is None or getattr(orig, '_wrapped', False):
    return

  def call(self, *a, **k):
    if _STACK and _STACK[-1] is self:
      return orig(self, *a, **k)
    if not self._setup_done:
      self._setup_done = True
      self.setup()
    # every call of a compact module re-binds the same submodule names (flax shares the
    # parameters of a module instance across calls: PropMLP is applied at two levels)
    self._counters = {}
    _STACK.append(self)
    try:
      return orig(self, *a, **k)
    finally:
      _STACK.pop()
  call._wrapped = True
  cls.__call__ = call


_orig_init_subclass = Module.__init_subclass__.__func__


def _init_subclass(cls, **kw):
  _wrap_call(cls)


Module.__init_subclass__ = classmethod(_init_subclass)
Module._run = lambda self, *a, **k: self(*a, **k)


class Dense(Module):
  features: int = 0
  kernel_init: object = None

  def __call__(self, x):
    from jax.numpy import _state
    dt = _np.float64 if _state['x64'] else _np.float32     # fp64 inside finite-difference probes
    w = _np.asarray(self._params['kernel'], dt)
    b = _np.asarray(self._params['bias'], dt)
    assert w.shape == (x.shape[-1], self.features), (self._name, w.shape, x.shape, self.features)
    return (_np.matmul(_np.asarray(x, dt), w) + b).astype(dt)


class Embed(Module):
  num_embeddings: int = 0
  features: int = 0
  embedding_init: object = None

  def __call__(self, idx):
    return _np.asarray(self._params['embedding'], _np.float32)[_np.asarray(idx)]


_wrap_call(Dense)
_wrap_call(Embed)
