"""Generate tests/golden/host.npz by EXECUTING the reference's own host-side modules
(internal/image.py, internal/raw_utils.py, the pose algebra of internal/camera_utils.py).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_host.py
Inputs are seeded and stored next to the outputs, so the tests need nothing but the .npz.
"""
import math
import os
import sys
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'standin'))
sys.path.insert(0, '/root/reference')
np.math = math
for missing in ['dm_pix', 'rawpy', 'mediapy', 'optax', 'pycolmap', 'matplotlib', 'tensorflow']:
  try:
    __import__(missing)
  except Exception:  # pylint: disable=broad-except
    sys.modules[missing] = mock.MagicMock()

import jax.numpy as jnp  # noqa: E402,F401  (the stand-in)
from internal import camera_utils, image, raw_utils  # noqa: E402


def random_poses(rng, n, spread=1.0):
  out = []
  for _ in range(n):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
      q[:, 0] *= -1
    out.append(np.concatenate([q, rng.normal(size=(3, 1)) * spread], axis=1))
  return np.stack(out)


def facing_poses(rng, n):
  """Cameras roughly on a ring looking at the origin (a 360 capture)."""
  out = []
  for i in range(n):
    a = 2 * np.pi * i / n + rng.normal() * 0.05
    eye = np.array([2.5 * np.cos(a), 2.5 * np.sin(a), 0.4 + 0.2 * rng.normal()])
    out.append(camera_utils.viewmatrix(eye, np.array([0., 0., 1.]), eye))
  return np.stack(out)


def main():
  rng = np.random.default_rng(123)
  out = {}
  # ---- image.py
  img = rng.uniform(0.1, 0.9, (24, 32, 3))
  ccm = np.eye(3) + rng.normal(size=(3, 3)) * 0.08
  ref = np.clip((img.reshape(-1, 3) @ ccm).reshape(img.shape) + 0.05 * img ** 2 - 0.03, 0, 1)
  out['cc_img'], out['cc_ref'] = img, ref
  out['cc_out'] = np.asarray(image.color_correct(img, ref))
  x = np.linspace(-0.5, 2.0, 501)
  out['srgb_x'] = x
  out['srgb_to_linear'] = np.asarray(image.srgb_to_linear(x, xnp=np))
  out['linear_to_srgb'] = np.asarray(image.linear_to_srgb(x, xnp=np))
  big = rng.uniform(size=(12, 8, 3))
  out['ds_in'], out['ds_out'] = big, np.asarray(image.downsample(big, 4))
  out['psnr_in'] = np.array([1e-4, 3e-3, 0.2])
  out['psnr_out'] = np.asarray(image.mse_to_psnr(out['psnr_in']))
  # ---- raw_utils.py
  bayer = rng.uniform(size=(10, 14))
  out['bayer'], out['demosaic'] = bayer, np.asarray(raw_utils.bilinear_demosaic(bayer, xnp=np))
  px, py = np.meshgrid(np.arange(5), np.arange(4), indexing='xy')
  out['mask_px'], out['mask_py'] = px, py
  out['bayer_mask'] = raw_utils.pixels_to_bayer_mask(px, py)
  raw = rng.uniform(size=(6, 7, 3))
  cam2rgb = np.eye(3) + rng.normal(size=(3, 3)) * 0.1
  out['pp_raw'], out['pp_cam2rgb'] = raw, cam2rgb
  out['pp_auto'] = np.asarray(raw_utils.postprocess_raw(raw, cam2rgb, None, xnp=np))
  out['pp_fixed'] = np.asarray(raw_utils.postprocess_raw(raw, cam2rgb, 0.6, xnp=np))
  exifs = []
  for i in range(3):
    exifs.append({'BlackLevel': 64 + i, 'WhiteLevel': 1023,
                  'AsShotNeutral': ' '.join(str(v) for v in rng.uniform(0.4, 1.0, 3)),
                  'ColorMatrix2': ' '.join(str(v) for v in (np.eye(3) + rng.normal(size=(3, 3)) * 0.2).ravel()),
                  'NoiseProfile': ' '.join(str(v) for v in rng.uniform(1e-4, 1e-3, 6)),
                  'ShutterSpeed': f'1/{[30, 120, 30][i]}'})
  meta = raw_utils.process_exif(exifs)
  out['exif_cam2rgb'] = meta['cam2rgb']
  out['exif_shutter'] = meta['ShutterSpeed']
  out['exif_json'] = np.array([repr(exifs)])
  est = rng.uniform(size=(9, 11, 3))
  gt = est * np.array([1.3, 0.8, 1.1]) + np.array([0.05, -0.02, 0.0]) + rng.normal(size=est.shape) * 0.01
  out['aff_est'], out['aff_gt'] = est, gt
  out['aff_out'] = np.asarray(raw_utils.match_images_affine(est, gt))
  # ---- camera_utils.py pose algebra
  poses = random_poses(rng, 9)
  out['poses'] = poses
  rp, rt = camera_utils.recenter_poses(poses)
  out['recenter_poses'], out['recenter_transform'] = rp, rt
  out['average_pose'] = camera_utils.average_pose(poses)
  ring = facing_poses(rng, 12)
  out['ring'] = ring
  out['focus_point'] = camera_utils.focus_point_fn(ring)
  pp, pt = camera_utils.transform_poses_pca(ring.copy())
  out['pca_poses'], out['pca_transform'] = pp, pt
  out['ellipse'] = camera_utils.generate_ellipse_path(pp, n_frames=10, z_variation=0.3, z_phase=0.25)
  out['ellipse_plain'] = camera_utils.generate_ellipse_path(pp, n_frames=7, const_speed=False)
  bounds = np.array([[1.2, 7.5], [1.0, 9.0], [1.5, 8.0]])
  out['bounds'] = bounds
  out['spiral'] = camera_utils.generate_spiral_path(rp, bounds, n_frames=8)
  out['interp_path'] = camera_utils.generate_interpolated_path(ring[:6], n_interp=4)
  out['interp_1d'] = np.asarray(camera_utils.interpolate_1d(np.log(np.array([1., 2., 1.5, 3., 2.5, 4., 3.])), 3, 5, 20))
  out['pad_poses'] = camera_utils.pad_poses(poses)
  np.savez_compressed(os.path.join(HERE, 'host.npz'), **out)
  print('wrote host.npz:', sorted(out))


if __name__ == '__main__':
  main()
