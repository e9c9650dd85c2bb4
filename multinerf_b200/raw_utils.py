"""RawNeRF pre- and post-processing around the path (reference: internal/raw_utils.py): Bayer masks for
the mosaic loss, bilinear demosaicking, EXIF -> colour pipeline metadata, raw dataset loading, the
minimal raw -> sRGB post-process, and the affine colour matching used by eval.py.

Host numpy: these run once per image at load / evaluation time.  Reading .dng files needs `rawpy`,
which this image does not ship; `load_raw_images` says so when called without it (everything else,
including `load_raw_dataset` on pre-extracted `.npy` Bayer planes, works without).
"""
import glob
import json
import os

import numpy as np

from . import image as lib_image
from . import utils


def postprocess_raw(raw, camtorgb, exposure=None):
  """Demosaicked camera-space raw -> sRGB: colour matrix, expose `exposure` to white (97th percentile
  when None), clip, sRGB curve (raw_utils.py:35-66)."""
  raw = np.asarray(raw)
  camtorgb = np.asarray(camtorgb)
  if raw.shape[-1] != 3:
    raise ValueError(f'raw.shape[-1] is {raw.shape[-1]}, expected 3')
  if camtorgb.shape != (3, 3):
    raise ValueError(f'camtorgb.shape is {camtorgb.shape}, expected (3, 3)')
  rgb_linear = np.matmul(raw, camtorgb.T)
  if exposure is None:
    exposure = np.percentile(rgb_linear, 97)
  return lib_image.linear_to_srgb(np.clip(rgb_linear / exposure, 0, 1))


def pixels_to_bayer_mask(pix_x, pix_y):
  """One-hot RGB mask of the RGGB mosaic at integer pixel coordinates (raw_utils.py:69-77)."""
  ex, ey = pix_x % 2 == 0, pix_y % 2 == 0
  r = ex & ey
  g = (~ex & ey) | (ex & ~ey)
  b = ~ex & ~ey
  return np.stack([r, g, b], -1).astype(np.float32)


def bilinear_demosaic(bayer):
  """[H, W] RGGB mosaic -> [H, W, 3] by bilinear interpolation (raw_utils.py:80-147).

  Red and blue: the measured samples are scattered into a full-resolution plane of zeros and spread with
  two cyclic separable passes (a sample keeps its value, a gap takes the mean of its two neighbours along
  the pass) -- the same arithmetic, pairing and wrap-around at the right/bottom edges as the reference's
  roll-based 2x upsampling.  Green: every missing site is the mean of its four cross neighbours."""
  bayer = np.asarray(bayer)
  H, W = bayer.shape

  def spread(plane):
    h = plane + .5 * (np.roll(plane, 1, axis=1) + np.roll(plane, -1, axis=1))
    return h + .5 * (np.roll(h, 1, axis=0) + np.roll(h, -1, axis=0))
  r = np.zeros_like(bayer)
  r[0::2, 0::2] = bayer[0::2, 0::2]
  b = np.zeros_like(bayer)
  b[1::2, 1::2] = bayer[1::2, 1::2]
  g = np.zeros_like(bayer)
  g[0::2, 1::2] = bayer[0::2, 1::2]
  g[1::2, 0::2] = bayer[1::2, 0::2]
  cross = .25 * np.roll(g, -1, axis=1)
  cross = cross + .25 * np.roll(g, 1, axis=1)
  cross = cross + .25 * np.roll(g, -1, axis=0)
  cross = cross + .25 * np.roll(g, 1, axis=0)
  measured = np.zeros((H, W), bool)
  measured[0::2, 1::2] = True
  measured[1::2, 0::2] = True
  green = np.where(measured, g, cross)
  return np.stack([spread(r), green, spread(b)], -1)


def load_raw_images(image_dir, image_names=None):
  """Raw Bayer planes + EXIF dicts (raw_utils.py:153-190).  Per image: `<name>.dng` (needs rawpy) or a
  pre-extracted `<name>.npy` Bayer plane, and `<name>.json` as written by `exiftool -json`."""
  if not os.path.exists(image_dir):
    raise ValueError(f'Raw image folder {image_dir} does not exist.')

  def load_one(image_name):
    base = os.path.join(image_dir, os.path.splitext(image_name)[0])
    if os.path.exists(base + '.npy'):
      raw = np.load(base + '.npy')
    else:
      try:
        import rawpy
      except ImportError as e:
        raise ImportError(f'reading {base}.dng needs the `rawpy` package (not in this image); '
                          'extract the Bayer plane to a .npy next to it instead') from e
      with open(base + '.dng', 'rb') as f:
        raw = rawpy.imread(f).raw_image
    with open(base + '.json', 'rb') as f:
      exif = json.load(f)[0]
    return raw, exif
  if image_names is None:
    found = sorted(glob.glob(os.path.join(image_dir, '*.dng')) + glob.glob(os.path.join(image_dir, '*.npy')))
    image_names = sorted({os.path.basename(f) for f in found}, key=lambda n: os.path.splitext(n)[0])
  raws, exifs = zip(*[load_one(x) for x in image_names])
  return np.stack(raws, axis=0).astype(np.float32), exifs


_PERCENTILE_LIST = (80, 90, 97, 99, 100)        # exposure sweep shown in the training logs
_EXIF_KEYS = ('BlackLevel', 'WhiteLevel', 'AsShotNeutral', 'ColorMatrix2', 'NoiseProfile')
# reference-illuminant XYZ <- linear sRGB (Lindbloom)
_RGB2XYZ = np.array([[0.4124564, 0.3575761, 0.1804375],
                     [0.2126729, 0.7151522, 0.0721750],
                     [0.0193339, 0.1191920, 0.9503041]])


def process_exif(exifs):
  """List of per-image EXIF dicts -> dict of arrays + the camera -> sRGB matrices (raw_utils.py:214-272):
  cam -> white-balanced cam (1 / AsShotNeutral) -> XYZ (ColorMatrix2, rows normalised) -> RGB."""
  meta = {}
  first = exifs[0]
  for key in _EXIF_KEYS:
    v = first.get(key)
    if v is None:
      continue
    if isinstance(v, (int, float)):
      vals = [x[key] for x in exifs]
    elif isinstance(v, str):
      vals = [[float(z) for z in x[key].split(' ')] for x in exifs]
    meta[key] = np.squeeze(np.array(vals))
  # shutter speeds are strings of the form '1/N'
  meta['ShutterSpeed'] = np.fromiter((1. / float(x['ShutterSpeed'].split('/')[1]) for x in exifs), float)
  whitebalance = meta['AsShotNeutral'].reshape(-1, 3)
  cam2camwb = np.array([np.diag(1. / x) for x in whitebalance])
  rgb2camwb = meta['ColorMatrix2'].reshape(-1, 3, 3) @ _RGB2XYZ
  rgb2camwb /= rgb2camwb.sum(axis=-1, keepdims=True)
  meta['cam2rgb'] = np.linalg.inv(rgb2camwb) @ cam2camwb
  return meta


def load_raw_dataset(split, data_dir, image_names, exposure_percentile, n_downsample):
  """RawNeRF inputs (raw_utils.py:275-385): demosaicked [N, H/n, W/n, 3] images in [0, 1] (+ noise),
  metadata with per-image exposure indices / relative shutter speeds and the `postprocess_fn`, and whether
  the scene is a "test scene" with an HDR+ ground-truth frame."""
  image_dir = os.path.join(data_dir, 'raw')
  testimg_file = os.path.join(data_dir, 'hdrplus_test/merged.dng')
  testimg_npy = os.path.join(data_dir, 'hdrplus_test/merged.npy')
  testscene = os.path.exists(testimg_file) or os.path.exists(testimg_npy)
  if testscene:
    image_dir = os.path.join(image_dir, split.value)
    image_names = None if split == utils.DataSplit.TEST else image_names[1:]
  raws, exifs = load_raw_images(image_dir, image_names)
  meta = process_exif(exifs)
  shutter_ratio = 1.
  if testscene and split == utils.DataSplit.TEST:
    if os.path.exists(testimg_npy):
      testraw = np.load(testimg_npy)
    else:
      import rawpy
      with open(testimg_file, 'rb') as f:
        testraw = rawpy.imread(f).raw_image
    testraw = testraw.astype(np.float32) / 4.            # HDR+ output carries 2 extra fixed-point bits
    shutter_ratio = meta['ShutterSpeed'][0] / meta['ShutterSpeed'][-1]
    raws = testraw[None]
    meta = {k: meta[k][:1] for k in meta}
  shutter_speeds = meta['ShutterSpeed']
  unique_shutters = np.sort(np.unique(shutter_speeds))[::-1]      # index 0 = slowest = brightest
  exposure_idx = np.zeros_like(shutter_speeds, dtype=np.int32)
  for i, s in enumerate(unique_shutters):
    exposure_idx[shutter_speeds == s] = i
  meta['exposure_idx'] = exposure_idx
  meta['unique_shutters'] = unique_shutters
  meta['exposure_values'] = shutter_speeds / unique_shutters[0]
  black = meta['BlackLevel'].reshape(-1, 1, 1)
  white = meta['WhiteLevel'].reshape(-1, 1, 1)
  images = (raws - black) / (white - black) * shutter_ratio
  image0_rgb = bilinear_demosaic(images[0]) @ meta['cam2rgb'][0].T
  exposure = np.percentile(image0_rgb, exposure_percentile)
  meta['exposure'] = exposure
  meta['exposure_levels'] = {p: np.percentile(image0_rgb, p) for p in _PERCENTILE_LIST}
  cam2rgb0 = meta['cam2rgb'][0]
  meta['postprocess_fn'] = lambda z, x=exposure: postprocess_raw(z, cam2rgb0, x)

  def process(x):
    x = bilinear_demosaic(x)
    return lib_image.downsample(x, n_downsample) if n_downsample > 1 else x
  images = np.stack([process(im) for im in images], axis=0)
  return images, meta, testscene


def best_fit_affine(x, y, axis):
  """Least-squares a, b with a * x + b = y (raw_utils.py:388-396)."""
  x_m, y_m = x.mean(axis=axis), y.mean(axis=axis)
  a = ((x * y).mean(axis=axis) - x_m * y_m) / ((x * x).mean(axis=axis) - x_m * x_m)
  return a, y_m - a * x_m


def match_images_affine(est, gt, axis=(0, 1)):
  """Fit gt -> est (robust to a noisy `est`), then map est back into gt's space (raw_utils.py:399-406)."""
  a, b = best_fit_affine(gt, est, axis=axis)
  return (est - b) / a
