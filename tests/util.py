"""Shared helpers for the tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden(name):
  return np.load(os.path.join(GOLDEN, name + '.npz'))


def T(x):
  return torch.as_tensor(np.asarray(x))


def close(a, b, atol=1e-5, rtol=1e-5, msg=''):
  a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
  b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
  assert a.shape == b.shape, (msg, a.shape, b.shape)
  a, b = np.atleast_1d(a), np.atleast_1d(b)
  both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
  both_nan = np.isnan(a) & np.isnan(b)
  err = np.abs(a - b)
  err[both_inf | both_nan] = 0
  tol = atol + rtol * np.abs(b)
  tol[both_inf | both_nan] = 1
  bad = ~(err <= tol)
  assert not bad.any(), (f'{msg}: {bad.sum()} / {bad.size} mismatches, max err '
                         f'{np.nanmax(err):.3e} at {np.unravel_index(np.nanargmax(err), err.shape)}')
