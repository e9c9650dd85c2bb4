"""Image-space helpers around the path (reference: internal/image.py): PSNR/SSIM conversions, sRGB
transfer functions, area downsampling, the quadratic colour-correction fit of eval.py, and the metric
harness.  Host numpy throughout (these run once per test image, not per ray).

`ssim` restates dm_pix.ssim (the dependency the reference calls, image.py:133; not vendored in
/root/reference): Gaussian window 11, sigma 1.5, k1 = .01, k2 = .03, max_val 1, 'valid' windows, mean
over pixels and channels.  PARITY: pinned against an independent scipy implementation in the tests, not
against dm_pix itself.
"""
import numpy as np

_EPS32 = float(np.finfo(np.float32).eps)


def mse_to_psnr(mse):
  """image.py:28-30 (maximum pixel value 1)."""
  return -10. / np.log(10.) * np.log(mse)


def psnr_to_mse(psnr):
  return np.exp(-0.1 * np.log(10.) * psnr)


def ssim_to_dssim(ssim):
  return (1 - ssim) / 2


def dssim_to_ssim(dssim):
  return 1 - 2 * dssim


def linear_to_srgb(linear, eps=None):
  """image.py:48-56."""
  eps = _EPS32 if eps is None else eps
  linear = np.asarray(linear)
  lo = 323 / 25 * linear
  hi = (211 * np.maximum(eps, linear) ** (5 / 12) - 11) / 200
  return np.where(linear <= 0.0031308, lo, hi)


def srgb_to_linear(srgb, eps=None):
  """image.py:59-67."""
  eps = _EPS32 if eps is None else eps
  srgb = np.asarray(srgb)
  lo = 25 / 323 * srgb
  hi = np.maximum(eps, (200 * srgb + 11) / 211) ** (12 / 5)
  return np.where(srgb <= 0.04045, lo, hi)


def downsample(img, factor):
  """Area downsample; the factor must divide both image sides (image.py:70-78)."""
  sh = img.shape
  if sh[0] % factor or sh[1] % factor:
    raise ValueError(f'Downsampling factor {factor} does not evenly divide image shape {sh[:2]}')
  return img.reshape((sh[0] // factor, factor, sh[1] // factor, factor) + sh[2:]).mean((1, 3))


def color_correct(img, ref, num_iters=5, eps=0.5 / 255):
  """Per-channel quadratic colour warp of `img` onto `ref`, refit `num_iters` times while the set of
  unsaturated pixels settles (image.py:81-124)."""
  if img.shape[-1] != ref.shape[-1]:
    raise ValueError(f'img\'s {img.shape[-1]} and ref\'s {ref.shape[-1]} channels must match')
  C = img.shape[-1]
  x = np.asarray(img, np.float64).reshape(-1, C)
  y = np.asarray(ref, np.float64).reshape(-1, C)
  inside = lambda z: (z >= eps) & (z <= 1 - eps)
  ok0 = inside(x)
  for _ in range(num_iters):
    # design matrix: upper-triangular quadratic terms, linear terms, bias
    cols = [x[:, c:c + 1] * x[:, c:] for c in range(C)] + [x, np.ones_like(x[:, :1])]
    A = np.concatenate(cols, axis=-1)
    warp = []
    for c in range(C):
      keep = ok0[:, c] & inside(x[:, c]) & inside(y[:, c])
      w = np.linalg.lstsq(np.where(keep[:, None], A, 0), np.where(keep, y[:, c], 0), rcond=-1)[0]
      assert np.all(np.isfinite(w))
      warp.append(w)
    x = np.clip(A @ np.stack(warp, axis=-1), 0, 1)
  return x.reshape(img.shape)


def _gauss_window(size, sigma):
  r = np.arange(size, dtype=np.float64) - (size - 1) / 2
  g = np.exp(-0.5 * (r / sigma) ** 2)
  return g / g.sum()


def _filt_valid(x, win):
  """Separable 'valid' correlation over the two leading (spatial) axes of [H, W, C]."""
  n = len(win)
  H, W = x.shape[:2]
  out = sum(win[i] * x[i:H - n + 1 + i] for i in range(n))
  return sum(win[i] * out[:, i:W - n + 1 + i] for i in range(n))


def ssim(a, b, max_val=1.0, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03):
  """Mean structural similarity of two [H, W, C] images (dm_pix.ssim defaults)."""
  a = np.asarray(a, np.float64)
  b = np.asarray(b, np.float64)
  if a.shape != b.shape or a.ndim != 3:
    raise ValueError(f'ssim expects two [H, W, C] images of one shape, got {a.shape} and {b.shape}')
  win = _gauss_window(filter_size, filter_sigma)
  mu_a, mu_b = _filt_valid(a, win), _filt_valid(b, win)
  saa = _filt_valid(a * a, win) - mu_a * mu_a
  sbb = _filt_valid(b * b, win) - mu_b * mu_b
  sab = _filt_valid(a * b, win) - mu_a * mu_b
  # dm_pix clips the (co)variances away from tiny negative round-off
  eps = np.finfo(np.float32).eps ** 2
  saa, sbb = np.maximum(eps, saa), np.maximum(eps, sbb)
  sab = np.sign(sab) * np.minimum(np.sqrt(saa * sbb), np.abs(sab))
  c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
  num = (2 * mu_a * mu_b + c1) * (2 * sab + c2)
  den = (mu_a ** 2 + mu_b ** 2 + c1) * (saa + sbb + c2)
  return float(np.mean(num / den))


class MetricHarness:
  """psnr + ssim of a predicted image against the ground truth (image.py:127-141)."""

  def __call__(self, rgb_pred, rgb_gt, name_fn=lambda s: s):
    rgb_pred = np.asarray(rgb_pred, np.float64)
    rgb_gt = np.asarray(rgb_gt, np.float64)
    psnr = float(mse_to_psnr(((rgb_pred - rgb_gt) ** 2).mean()))
    return {name_fn('psnr'): psnr, name_fn('ssim'): ssim(rgb_pred, rgb_gt)}
