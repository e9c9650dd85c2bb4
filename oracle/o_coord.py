"""Oracle (test infrastructure): coordinate warps and encodings, torch-CPU.

Follows /root/reference/internal/coord.py:
  contract :21-27   track_linearize :39-60 (closed-form Jacobian instead of
  jax.linearize; checked against autograd in tests)   construct_ray_warps :63-99
  expected_sin :102-104   integrated_pos_enc :107-126
  lift_and_diagonalize :129-133   pos_enc :136-147
and /root/reference/internal/ref_utils.py:
  reflect :22-37   l2_normalize :40-42   generate_ide_fn :98-159
"""
import math

import numpy as np
import torch

from . import o_math

EPS = o_math.EPS


def contract(x):
  x_mag_sq = torch.clamp((x ** 2).sum(dim=-1, keepdim=True), min=EPS)
  return torch.where(x_mag_sq <= 1, x, ((2 * torch.sqrt(x_mag_sq) - 1) / x_mag_sq) * x)


def contract_jacobian(x):
  """d contract / dx as a [..., 3, 3] matrix.

  r2 = max(eps, |x|^2); inside the unit ball J = I; outside, with r = sqrt(r2),
  J = (2/r - 1/r2) I + (2/r2^2 - 2/(r2 r)) x x^T.  (The eps clamp only matters at
  |x|^2 < eps where the branch is the identity anyway.)
  """
  r2 = torch.clamp((x ** 2).sum(dim=-1, keepdim=True), min=EPS)[..., None]
  r = torch.sqrt(r2)
  eye = torch.eye(x.shape[-1], dtype=x.dtype)
  s = 2 / r - 1 / r2
  c = 2 / (r2 * r2) - 2 / (r2 * r)
  outer = x[..., :, None] * x[..., None, :]
  return torch.where(r2 <= 1, eye.expand_as(outer), s * eye + c * outer)


def track_linearize_contract(mean, cov):
  """coord.track_linearize(coord.contract, mean, cov): (contract(mean), J cov J^T)."""
  if mean.dim() + 1 != cov.dim():
    raise ValueError('cov must be non-diagonal')
  jac = contract_jacobian(mean)
  return contract(mean), jac @ cov @ jac.transpose(-1, -2)


def track_linearize_autograd(fn, mean, cov):
  """Generic version via torch.func (used only by tests to check the closed form)."""
  from torch.func import jacrev, vmap
  flat = mean.reshape(-1, mean.shape[-1])
  jac = vmap(jacrev(fn))(flat).reshape(mean.shape + (mean.shape[-1],))
  return fn(mean), jac @ cov @ jac.transpose(-1, -2)


def construct_ray_warps(fn, t_near, t_far):
  """coord.py:63-99; `fn` is None, 'piecewise' or one of the names below."""
  if fn is None:
    fwd = inv = (lambda x: x)
  elif fn == 'piecewise':
    fwd = lambda x: torch.where(x < 1, .5 * x, 1 - .5 / x)
    inv = lambda x: torch.where(x < .5, 2 * x, .5 / (1 - x))
  else:
    table = {
        'reciprocal': (torch.reciprocal, torch.reciprocal),
        'log': (torch.log, torch.exp),
        'exp': (torch.exp, torch.log),
        'sqrt': (torch.sqrt, torch.square),
        'square': (torch.square, torch.sqrt),
    }
    fwd, inv = table[fn]
  s_near, s_far = fwd(t_near), fwd(t_far)
  t_to_s = lambda t: (fwd(t) - s_near) / (s_far - s_near)
  s_to_t = lambda s: inv(s * s_far + (1 - s) * s_near)
  return t_to_s, s_to_t


def expected_sin(mean, var):
  return torch.exp(-0.5 * var) * o_math.safe_sin(mean)


def integrated_pos_enc(mean, var, min_deg, max_deg):
  scales = 2.0 ** torch.arange(min_deg, max_deg, dtype=mean.dtype)
  shape = mean.shape[:-1] + (-1,)
  scaled_mean = (mean[..., None, :] * scales[:, None]).reshape(shape)
  scaled_var = (var[..., None, :] * scales[:, None] ** 2).reshape(shape)
  return expected_sin(
      torch.cat([scaled_mean, scaled_mean + 0.5 * math.pi], dim=-1),
      torch.cat([scaled_var] * 2, dim=-1))


def lift_and_diagonalize(mean, cov, basis):
  """basis: [3, K].  True-fp32 matmuls (math.matmul forces HIGHEST, math.py:21-23)."""
  fn_mean = mean @ basis
  fn_cov_diag = (basis * (cov @ basis)).sum(dim=-2)
  return fn_mean, fn_cov_diag


def pos_enc(x, min_deg, max_deg, append_identity=True):
  scales = 2.0 ** torch.arange(min_deg, max_deg, dtype=x.dtype)
  shape = x.shape[:-1] + (-1,)
  scaled_x = (x[..., None, :] * scales[:, None]).reshape(shape)
  four_feat = torch.sin(torch.cat([scaled_x, scaled_x + 0.5 * math.pi], dim=-1))
  return torch.cat([x, four_feat], dim=-1) if append_identity else four_feat


# ---------------------------------------------------------------- ref_utils

def reflect(viewdirs, normals):
  return 2.0 * (normals * viewdirs).sum(dim=-1, keepdim=True) * normals - viewdirs


def l2_normalize(x, eps=EPS):
  return x / torch.sqrt(torch.clamp((x ** 2).sum(dim=-1, keepdim=True), min=eps))


def _gen_binom(a, k):
  return float(np.prod(a - np.arange(k))) / math.factorial(k)


def _assoc_legendre_coeff(l, m, k):
  return ((-1) ** m * 2 ** l * math.factorial(l) / math.factorial(k) /
          math.factorial(l - k - m) * _gen_binom(0.5 * (l + k + m - 1.0), l))


def _sph_harm_coeff(l, m, k):
  return (math.sqrt((2.0 * l + 1.0) * math.factorial(l - m) /
                    (4.0 * math.pi * math.factorial(l + m))) * _assoc_legendre_coeff(l, m, k))


def ide_tables(deg_view):
  """(ml_array [2, n] int, mat [l_max+1, n] float64) of ref_utils.py:84-123."""
  if deg_view > 5:
    raise ValueError('Only deg_view of at most 5 is numerically stable.')
  ml = [(m, 2 ** i) for i in range(deg_view) for m in range(2 ** i + 1)]
  ml_array = np.array(ml).T
  l_max = 2 ** (deg_view - 1)
  mat = np.zeros((l_max + 1, ml_array.shape[1]))
  for i, (m, l) in enumerate(ml_array.T):
    for k in range(l - m + 1):
      mat[k, i] = _sph_harm_coeff(int(l), int(m), k)
  return ml_array, mat


def generate_ide_fn(deg_view):
  """Integrated directional encoding (ref_utils.py:98-159), complex64 as in the ref."""
  ml_array, mat_np = ide_tables(deg_view)

  def ide(xyz, kappa_inv):
    mat = torch.tensor(mat_np, dtype=xyz.dtype)
    x, y, z = xyz[..., 0:1], xyz[..., 1:2], xyz[..., 2:3]
    vmz = torch.cat([z ** i for i in range(mat.shape[0])], dim=-1)
    xy = torch.complex(x, y)
    vmxy = torch.cat([xy ** int(m) for m in ml_array[0, :]], dim=-1)
    sph = vmxy * (vmz @ mat)
    sigma = torch.tensor(0.5 * ml_array[1, :] * (ml_array[1, :] + 1), dtype=xyz.dtype)
    out = sph * torch.exp(-sigma * kappa_inv)
    return torch.cat([out.real, out.imag], dim=-1)

  return ide
