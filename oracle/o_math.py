"""Oracle (test infrastructure): scalar helpers of the hot path, torch-CPU.

Follows /root/reference/internal/math.py:
  safe_sin / safe_cos      math.py:26-38
  safe_exp (+ custom JVP)  math.py:41-54
  log_lerp, learning_rate_decay   math.py:57-98
  interp (gather based)    math.py:101-105
  sorted_interp            math.py:108-127
and /root/reference/internal/image.py:28-30 (mse_to_psnr), :48-56 (linear_to_srgb).
"""
import math

import torch

EPS = float(torch.finfo(torch.float32).eps)  # 2**-23, the reference's jnp.finfo(float32).eps
_T_SAFE = 100.0 * math.pi


def _safe_arg(x):
  # math.py:26-28: where(|x| < 100*pi, x, x % (100*pi)); `%` is Python-style
  # (sign of the divisor) for jnp and for torch.remainder alike.
  t = torch.tensor(_T_SAFE, dtype=x.dtype)
  return torch.where(x.abs() < t, x, torch.remainder(x, t))


def safe_sin(x):
  return torch.sin(_safe_arg(x))


def safe_cos(x):
  return torch.cos(_safe_arg(x))


class _SafeExp(torch.autograd.Function):
  """exp(min(x, 88)) whose derivative is its own value (math.py:41-54)."""

  @staticmethod
  def forward(ctx, x):
    y = torch.exp(torch.clamp(x, max=88.0))
    ctx.save_for_backward(y)
    return y

  @staticmethod
  def backward(ctx, g):
    (y,) = ctx.saved_tensors
    return g * y


def safe_exp(x):
  return _SafeExp.apply(x)


def log_lerp(t, v0, v1):
  if v0 <= 0 or v1 <= 0:
    raise ValueError(f'Interpolants {v0} and {v1} must be positive.')
  lv0, lv1 = math.log(v0), math.log(v1)
  return math.exp(min(max(t, 0.0), 1.0) * (lv1 - lv0) + lv0)


def learning_rate_decay(step, lr_init, lr_final, max_steps, lr_delay_steps=0,
                        lr_delay_mult=1.0):
  """math.py:66-98, evaluated in Python floats (host side scalar)."""
  if lr_delay_steps > 0:
    delay_rate = lr_delay_mult + (1 - lr_delay_mult) * math.sin(
        0.5 * math.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
  else:
    delay_rate = 1.0
  return delay_rate * log_lerp(step / max_steps, lr_init, lr_final)


def interval_index(x, xp):
  """idx0 = #{xp <= x} - 1 clipped to [0, P]; the integer the resampler implies.

  Because xp is non-decreasing the mask `x >= xp` of math.py:113 is a prefix, so
  the "last True" position is this count minus one.
  x: [..., S], xp: [..., P+1]  ->  int64 [..., S]
  """
  mask = x[..., None, :] >= xp[..., :, None]           # [..., P+1, S]
  return mask.sum(dim=-2) - 1


def sorted_interp(x, xp, fp, return_index=False):
  """math.py:108-127 restated in index form (bit-identical to the mask form)."""
  n = xp.shape[-1]
  cnt = (x[..., None, :] >= xp[..., :, None]).sum(dim=-2)   # in [0, n]
  i0 = (cnt - 1).clamp(min=0)      # max(where(mask, v, v[0]))  -> v[0] when no True
  i1 = cnt.clamp(max=n - 1)        # min(where(~mask, v, v[-1])) -> v[-1] when all True
  xp0 = torch.gather(xp.expand(*x.shape[:-1], n), -1, i0)
  xp1 = torch.gather(xp.expand(*x.shape[:-1], n), -1, i1)
  fp0 = torch.gather(fp.expand(*x.shape[:-1], n), -1, i0)
  fp1 = torch.gather(fp.expand(*x.shape[:-1], n), -1, i1)
  off = torch.nan_to_num((x - xp0) / (xp1 - xp0), nan=0.0, posinf=float('inf'),
                         neginf=float('-inf'))
  # jnp.nan_to_num(z, 0) also maps +-inf to the largest finite values; the clip
  # that follows makes both conventions agree.
  off = off.clamp(0.0, 1.0)
  ret = fp0 + off * (fp1 - fp0)
  if return_index:
    return ret, cnt - 1
  return ret


def interp(x, xp, fp):
  """np.interp semantics per row (math.py:101-105, the gather-based variant)."""
  n = xp.shape[-1]
  xpe = xp.expand(*x.shape[:-1], n).contiguous()
  fpe = fp.expand(*x.shape[:-1], n).contiguous()
  idx = torch.searchsorted(xpe, x.contiguous(), right=True).clamp(1, n - 1)
  x0 = torch.gather(xpe, -1, idx - 1)
  x1 = torch.gather(xpe, -1, idx)
  f0 = torch.gather(fpe, -1, idx - 1)
  f1 = torch.gather(fpe, -1, idx)
  dx = x1 - x0
  slope = torch.where(dx.abs() <= torch.finfo(x.dtype).tiny, torch.zeros_like(dx),
                      (f1 - f0) / torch.where(dx == 0, torch.ones_like(dx), dx))
  out = f0 + slope * (x - x0)
  out = torch.where(x <= xpe[..., :1], fpe[..., :1], out)
  out = torch.where(x >= xpe[..., -1:], fpe[..., -1:], out)
  return out


def mse_to_psnr(mse):
  return -10.0 / math.log(10.0) * torch.log(mse)


def linear_to_srgb(linear):
  """image.py:48-56."""
  srgb0 = 323.0 / 25.0 * linear
  srgb1 = (211.0 * torch.clamp(linear, min=EPS) ** (5.0 / 12.0) - 11.0) / 200.0
  return torch.where(linear <= 0.0031308, srgb0, srgb1)
