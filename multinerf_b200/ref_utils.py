"""Host-side constants of the integrated directional encoding (Ref-NeRF).

Reference: internal/ref_utils.py:53-123 (`generalized_binomial_coeff`, `assoc_legendre_coeff`,
`sph_harm_coeff`, `get_ml_array`, the `mat` table built inside `generate_ide_fn`).  The device
kernels (csrc/refnerf.cu) evaluate  ide_i = (x+iy)^{m_i} * (sum_k mat[k,i] z^k) * exp(-sigma_i/kappa)
with sigma_i = l_i (l_i + 1) / 2, real parts first then imaginary parts.
"""
import math

import numpy as np


def _gen_binom(a, k):
  out = 1.0
  for j in range(k):
    out *= (a - j)
  return out / math.factorial(k)


def _legendre_coeff(l, m, k):
  # coefficient of cos^k(theta) sin^m(theta) in P_l^m(cos(theta))
  return ((-1) ** m * 2 ** l * math.factorial(l) / math.factorial(k) / math.factorial(l - k - m) *
          _gen_binom(0.5 * (l + k + m - 1.0), l))


def _sph_coeff(l, m, k):
  return math.sqrt((2.0 * l + 1.0) * math.factorial(l - m) / (4.0 * math.pi * math.factorial(l + m))) * \
      _legendre_coeff(l, m, k)


def ide_tables(deg_view):
  """Returns (m[n], l[n], mat[l_max+1, n] float64) for the n = sum_i (2^i + 1) (m, l) pairs."""
  if deg_view > 5:
    raise ValueError('Only deg_view of at most 5 is numerically stable.')
  ms, ls = [], []
  for i in range(deg_view):
    l = 2 ** i
    for m in range(l + 1):
      ms.append(m)
      ls.append(l)
  l_max = 2 ** (deg_view - 1)
  mat = np.zeros((l_max + 1, len(ms)))
  for i, (m, l) in enumerate(zip(ms, ls)):
    for k in range(l - m + 1):
      mat[k, i] = _sph_coeff(l, m, k)
  return np.array(ms, np.int32), np.array(ls, np.int32), mat


def ide_dim(deg_view):
  return 2 * sum(2 ** i + 1 for i in range(deg_view))
