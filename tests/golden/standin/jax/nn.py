import numpy as _np
from scipy import special as _sp

from .numpy import _cast


def softmax(x, axis=-1):
  return _cast(_sp.softmax(x, axis=axis))


def relu(x):
  return _np.maximum(x, 0)


def softplus(x):
  return _cast(_np.logaddexp(x, 0))


def sigmoid(x):
  return _cast(_sp.expit(x))


def silu(x):
  return x * sigmoid(x)


class _Init:
  def __getattr__(self, name):
    return lambda *a, **k: name


initializers = _Init()
