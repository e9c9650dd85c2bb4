"""Explicit-stream stub: a 'key' is a list of pre-drawn arrays consumed in call order."""


class Stream:
  """Holds named draws; `split` hands the same stream back so call order is preserved."""

  def __init__(self, draws):
    self.draws = list(draws)

  def pop(self, shape):
    import numpy as np
    d = np.asarray(self.draws.pop(0), dtype=np.float32)
    assert tuple(d.shape) == tuple(shape), (d.shape, shape)
    return d


def split(key, num=2):
  return tuple([key] * num)


def uniform(key, shape=(), minval=0.0, maxval=1.0):
  import numpy as np
  return (key.pop(shape) * np.float32(maxval - minval) + np.float32(minval)).astype(np.float32)


def normal(key, shape=()):
  return key.pop(shape)


def PRNGKey(seed):
  return Stream([])


def permutation(key, n):
  import numpy as np
  return np.arange(n)
