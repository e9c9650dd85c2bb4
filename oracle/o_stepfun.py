"""Oracle (test infrastructure): step-function algebra, torch-CPU.

Follows /root/reference/internal/stepfun.py (last axis = samples, `t` are bin
endpoints, `w` bin weights, `p` bin densities -- stepfun.py:15-23):
  searchsorted :30-53     inner_outer :64-77      lossfun_outer :80-86
  weight_to_pdf :89-91    pdf_to_weight :94-96    max_dilate :99-113
  max_dilate_weights :116-128   integrate_weights :131-150
  invert_cdf :153-161     sample :164-211         sample_intervals :214-263
  lossfun_distortion :266-276   weighted_percentile :298-308

Randomness is an explicit input (`jitter` = raw U[0,1) numbers) because the
reference's threefry stream (stepfun.py:209) cannot be reproduced without JAX.
"""
import torch

from . import o_math

EPS = o_math.EPS


def searchsorted(a, v):
  """(idx_lo, idx_hi) with a[idx_lo] <= v < a[idx_hi] (stepfun.py:30-53)."""
  n = a.shape[-1]
  i = torch.arange(n)
  v_ge_a = v[..., None, :] >= a[..., :, None]                      # [..., n, nv]
  idx_lo = torch.where(v_ge_a, i[:, None], i[:1, None]).amax(dim=-2)
  idx_hi = torch.where(~v_ge_a, i[:, None], i[-1:, None]).amin(dim=-2)
  return idx_lo, idx_hi


def inner_outer(t0, t1, y1):
  """Inner/outer measures of (t1, y1) on the intervals of t0 (stepfun.py:64-77)."""
  cy1 = torch.cat([torch.zeros_like(y1[..., :1]), torch.cumsum(y1, dim=-1)], dim=-1)
  idx_lo, idx_hi = searchsorted(t1, t0)
  cy1_lo = torch.gather(cy1, -1, idx_lo)
  cy1_hi = torch.gather(cy1, -1, idx_hi)
  y0_outer = cy1_hi[..., 1:] - cy1_lo[..., :-1]
  y0_inner = torch.where(idx_hi[..., :-1] <= idx_lo[..., 1:],
                         cy1_lo[..., 1:] - cy1_hi[..., :-1],
                         torch.zeros_like(y0_outer))
  return y0_inner, y0_outer


def lossfun_outer(t, w, t_env, w_env, eps=EPS):
  """stepfun.py:80-86."""
  _, w_outer = inner_outer(t, t_env, w_env)
  return torch.clamp(w - w_outer, min=0.0) ** 2 / (w + eps)


def weight_to_pdf(t, w, eps=EPS ** 2):
  return w / torch.clamp(t[..., 1:] - t[..., :-1], min=eps)


def pdf_to_weight(t, p):
  return p * (t[..., 1:] - t[..., :-1])


def max_dilate(t, w, dilation, domain=(-float('inf'), float('inf'))):
  """Max-pool dilation of a step function (stepfun.py:99-113)."""
  t0 = t[..., :-1] - dilation
  t1 = t[..., 1:] + dilation
  t_dilate, _ = torch.sort(torch.cat([t, t0, t1], dim=-1), dim=-1)
  t_dilate = t_dilate.clamp(domain[0], domain[1])
  covered = (t0[..., None, :] <= t_dilate[..., None]) & (t1[..., None, :] > t_dilate[..., None])
  w_dilate = torch.where(covered, w[..., None, :], torch.zeros((), dtype=w.dtype)).amax(dim=-1)
  return t_dilate, w_dilate[..., :-1]


def max_dilate_weights(t, w, dilation, domain=(-float('inf'), float('inf')),
                       renormalize=False, eps=EPS ** 2):
  """stepfun.py:116-128."""
  p = weight_to_pdf(t, w)
  t_dilate, p_dilate = max_dilate(t, p, dilation, domain=domain)
  w_dilate = pdf_to_weight(t_dilate, p_dilate)
  if renormalize:
    w_dilate = w_dilate / torch.clamp(w_dilate.sum(dim=-1, keepdim=True), min=eps)
  return t_dilate, w_dilate


def integrate_weights(w):
  """[0, min(1, cumsum(w[:-1])), 1]  (stepfun.py:131-150)."""
  cw = torch.clamp(torch.cumsum(w[..., :-1], dim=-1), max=1.0)
  shape = cw.shape[:-1] + (1,)
  return torch.cat([torch.zeros(shape, dtype=w.dtype), cw, torch.ones(shape, dtype=w.dtype)],
                   dim=-1)


def invert_cdf(u, t, w_logits, use_gpu_resampling=False, return_index=False):
  """stepfun.py:153-161."""
  w = torch.softmax(w_logits, dim=-1)
  cw = integrate_weights(w)
  if use_gpu_resampling:
    t_new = o_math.interp(u, cw, t)
    idx = o_math.interval_index(u, cw)
  else:
    t_new, idx = o_math.sorted_interp(u, cw, t, return_index=True)
  if return_index:
    return t_new, idx, cw
  return t_new


def sample_u(batch_shape, num_samples, jitter=None, single_jitter=False,
             deterministic_center=False, dtype=torch.float32):
  """The u grid of stepfun.py:190-209.  `jitter` = raw U[0,1) numbers of shape
  batch_shape + (1 if single_jitter else num_samples,), or None (rng=None)."""
  eps = EPS
  if jitter is None:
    if deterministic_center:
      pad = 1 / (2 * num_samples)
      u = torch.linspace(pad, 1.0 - pad - eps, num_samples, dtype=dtype)
    else:
      u = torch.linspace(0, 1.0 - eps, num_samples, dtype=dtype)
    u = u.expand(*batch_shape, num_samples)
  else:
    u_max = eps + (1 - eps) / num_samples
    max_jitter = (1 - u_max) / (num_samples - 1) - eps
    d = 1 if single_jitter else num_samples
    assert jitter.shape[-1] == d, (jitter.shape, d)
    u = torch.linspace(0, 1 - u_max, num_samples, dtype=dtype) + jitter.to(dtype) * max_jitter
  return u


def sample(jitter, t, w_logits, num_samples, single_jitter=False,
           deterministic_center=False, use_gpu_resampling=False, return_index=False):
  """stepfun.py:164-211 with explicit randomness."""
  u = sample_u(t.shape[:-1], num_samples, jitter, single_jitter, deterministic_center,
               dtype=t.dtype)
  return invert_cdf(u, t, w_logits, use_gpu_resampling=use_gpu_resampling,
                    return_index=return_index)


def sample_intervals(jitter, t, w_logits, num_samples, single_jitter=False,
                     domain=(-float('inf'), float('inf')), use_gpu_resampling=False,
                     return_index=False):
  """stepfun.py:214-263."""
  if num_samples <= 1:
    raise ValueError(f'num_samples must be > 1, is {num_samples}.')
  out = sample(jitter, t, w_logits, num_samples, single_jitter, deterministic_center=True,
               use_gpu_resampling=use_gpu_resampling, return_index=return_index)
  centers = out[0] if return_index else out
  mid = (centers[..., 1:] + centers[..., :-1]) / 2
  minval, maxval = domain
  first = torch.clamp(2 * centers[..., :1] - mid[..., :1], min=minval)
  last = torch.clamp(2 * centers[..., -1:] - mid[..., -1:], max=maxval)
  t_samples = torch.cat([first, mid, last], dim=-1)
  if return_index:
    return t_samples, out[1], out[2]
  return t_samples


def lossfun_distortion(t, w):
  """stepfun.py:266-276 (the O(S^2) form, as in the reference)."""
  ut = (t[..., 1:] + t[..., :-1]) / 2
  dut = (ut[..., :, None] - ut[..., None, :]).abs()
  loss_inter = (w * (w[..., None, :] * dut).sum(dim=-1)).sum(dim=-1)
  loss_intra = (w ** 2 * (t[..., 1:] - t[..., :-1])).sum(dim=-1) / 3
  return loss_inter + loss_intra


def weighted_percentile(t, w, ps):
  """stepfun.py:298-308: np.interp(ps/100, integrate_weights(w), t) per row."""
  cw = integrate_weights(w)
  q = (torch.tensor(ps, dtype=t.dtype) / 100).expand(*t.shape[:-1], len(ps))
  return o_math.interp(q, cw, t)
