"""Visualisations written to the training / evaluation summaries (reference: internal/vis.py).

Host numpy on finished [H, W, ...] renderings.  Implemented: weighted percentiles, the checker
matte, colour-mapped depth (`turbo` through its published polynomial fit -- matplotlib is not in this
image), the three-percentile depth triplet, coordinate-modulo maps, normals / roughness mattes and
`visualize_suite`.  Not implemented: the per-ray bundle plots (`visualize_rays`), which only feed
TensorBoard images.
"""
import numpy as np

_EPS = float(np.finfo(np.float32).eps)


def _np(x):
  return x.detach().cpu().numpy() if hasattr(x, 'detach') else np.asarray(x)


def weighted_percentile(x, w, ps, assume_sorted=False):
  """Weighted percentile(s) of one vector (vis.py:22-30)."""
  x, w = np.reshape(x, [-1]), np.reshape(w, [-1])
  if not assume_sorted:
    order = np.argsort(x)
    # the depth triplet passes a 3-channel value with a 1-channel weight: jnp's gather clamps the
    # out-of-range indices (vis.py:27-28), which is mirrored here
    x, w = x[order], w[np.minimum(order, w.shape[0] - 1)]
  acc_w = np.cumsum(w)
  return np.interp(np.array(ps) * (acc_w[-1] / 100), acc_w, x)


def sinebow(h):
  """Cyclic, uniform colormap (vis.py:33-36)."""
  f = lambda x: np.sin(np.pi * x) ** 2
  return np.stack([f(3 / 6 - h), f(5 / 6 - h), f(7 / 6 - h)], -1)


def turbo(x):
  """Google's turbo colormap on [0, 1] (the published degree-5 polynomial approximation)."""
  x = np.clip(np.asarray(x, np.float64), 0, 1)
  v = np.stack([np.ones_like(x), x, x ** 2, x ** 3, x ** 4, x ** 5], -1)
  r = v @ np.array([0.13572138, 4.61539260, -42.66032258, 132.13108234, -152.94239396, 59.28637943])
  g = v @ np.array([0.09140261, 2.19418839, 4.84296658, -14.18503333, 4.27729857, 2.82956604])
  b = v @ np.array([0.10667330, 12.64194608, -60.58204836, 110.36276771, -89.90310912, 27.34824973])
  return np.clip(np.stack([r, g, b], -1), 0, 1)


def gray(x):
  x = np.clip(np.asarray(x, np.float64), 0, 1)
  return np.stack([x, x, x], -1)


def matte(vis, acc, dark=0.8, light=1.0, width=8):
  """Composite over a checkerboard where nothing accumulated (vis.py:39-46)."""
  rows = (np.arange(acc.shape[0]) % (2 * width) // width)[:, None]
  cols = (np.arange(acc.shape[1]) % (2 * width) // width)[None, :]
  bg = np.where(np.logical_xor(rows, cols), light, dark)
  return vis * acc[:, :, None] + (bg * (1 - acc))[:, :, None]


def visualize_cmap(value, weight, colormap, lo=None, hi=None, percentile=99., curve_fn=lambda x: x,
                   modulus=None, matte_background=True):
  """Scalar (or 3-channel) map -> colours between weighted-percentile bounds (vis.py:49-110)."""
  lo_auto, hi_auto = weighted_percentile(value, weight, [50 - percentile / 2, 50 + percentile / 2])
  lo = lo or (lo_auto - _EPS)
  hi = hi or (hi_auto + _EPS)
  value, lo, hi = [curve_fn(x) for x in [value, lo, hi]]
  if modulus:
    value = np.mod(value, modulus) / modulus
  else:
    value = np.nan_to_num(np.clip((value - np.minimum(lo, hi)) / np.abs(hi - lo), 0, 1))
  if colormap:
    colorized = colormap(value)[:, :, :3]
  else:
    if value.ndim != 3 or value.shape[-1] != 3:
      raise ValueError(f'value must be [H, W, 3] without a colormap, got {value.shape}')
    colorized = value
  return matte(colorized, weight) if matte_background else colorized


def visualize_coord_mod(coords, acc):
  """Position of each surface point inside its unit cell (vis.py:113-115)."""
  return matte(((coords + 1) % 2) / 2, acc)


def visualize_suite(rendering, rays):
  """The image set train.py / eval.py log per test view (vis.py:188-260, without the ray-bundle plots)."""
  rendering = {k: (_np(v) if not isinstance(v, (list, tuple)) else v) for k, v in rendering.items()}
  depth_curve = lambda x: -np.log(x + _EPS)
  rgb, acc = rendering['rgb'], rendering['acc']
  d_mean, d_med = rendering['distance_mean'], rendering['distance_median']
  d_p5, d_p95 = rendering['distance_percentile_5'], rendering['distance_percentile_95']
  acc = np.where(np.isnan(d_mean), np.zeros_like(acc), acc)
  coords = _np(rays.origins) + _np(rays.directions) * d_mean[:, :, None]
  vis = {
      'color': rgb,
      'acc': acc,
      'color_matte': matte(rgb, acc),
      'depth_mean': visualize_cmap(d_mean, acc, turbo, curve_fn=depth_curve),
      'depth_median': visualize_cmap(d_med, acc, turbo, curve_fn=depth_curve),
      'depth_triplet': visualize_cmap(np.stack([2 * d_med - d_p5, d_med, d_p95], axis=-1), acc, None,
                                      curve_fn=lambda x: np.log(x + _EPS)),
      'coords_mod': visualize_coord_mod(coords, acc),
  }
  if 'rgb_cc' in rendering:
    vis['color_corrected'] = rendering['rgb_cc']
  for key, val in rendering.items():
    if key.startswith('normals') and not isinstance(val, (list, tuple)):
      vis[key] = matte(val / 2. + 0.5, acc)
  if 'roughness' in rendering:
    vis['roughness'] = matte(np.tanh(rendering['roughness']), acc)
  return vis
