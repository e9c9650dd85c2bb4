"""Generate tests/golden/camera.npz by EXECUTING the reference's own internal/camera_utils.py.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_camera.py
`pixels_to_rays` / `cast_ray_batch` / `convert_to_ndc` take an `xnp` module: the fixture holds the
outputs of both the float64 numpy path (`xnp=np`, what the reference's Dataset thread runs) and
the float32 path (`xnp=jnp` under the numpy stand-in for jax.numpy, what the reference runs inside
the train step when `cast_rays_in_train_step=True`).  Inputs are seeded and stored next to the
outputs, so the tests need nothing but the .npz.
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
for missing in ['dm_pix', 'cv2', 'rawpy', 'mediapy', 'optax', 'pycolmap', 'matplotlib', 'tensorflow',
                'scipy.interpolate', 'PIL', 'PIL.Image', 'PIL.ExifTags']:
  try:
    __import__(missing)
  except Exception:  # pylint: disable=broad-except
    sys.modules[missing] = mock.MagicMock()

import jax.numpy as jnp  # noqa: E402  (the stand-in)
from internal import camera_utils, utils  # noqa: E402

F = np.float32


def random_pose(rng):
  a = rng.normal(size=(3, 3))
  q, _ = np.linalg.qr(a)
  if np.linalg.det(q) < 0:
    q[:, 0] *= -1
  t = rng.uniform(-1.5, 1.5, (3, 1))
  return np.concatenate([q, t], axis=1)


def main():
  rng = np.random.default_rng(11)
  out = {}
  n_cam, B = 5, 257
  W, H = 160, 120
  focals = rng.uniform(100.0, 200.0, n_cam)
  pixtocams = np.stack([camera_utils.get_pixtocam(f, W, H) for f in focals])
  camtoworlds = np.stack([random_pose(rng) for _ in range(n_cam)])
  # forward-facing poses for the NDC case: small rotations about identity, looking down -z
  ndc_poses = []
  for _ in range(n_cam):
    w = rng.normal(size=3) * 0.08
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    R = np.eye(3) + K + 0.5 * K @ K
    q, _ = np.linalg.qr(R)
    q = q * np.sign(np.diag(q))[None, :]
    ndc_poses.append(np.concatenate([q, rng.uniform(-0.3, 0.3, (3, 1))], axis=1))
  ndc_poses = np.stack(ndc_poses)
  pix_x = rng.integers(0, W, B).astype(np.int32)
  pix_y = rng.integers(0, H, B).astype(np.int32)
  cam_idx = rng.integers(0, n_cam, (B, 1)).astype(np.int32)
  dist = dict(k1=0.05, k2=-0.02, k3=0.004, k4=0.0, p1=0.001, p2=-0.0015)
  pixtocam_ndc = camera_utils.get_pixtocam(150.0, W, H)
  out.update(pixtocams=pixtocams, camtoworlds=camtoworlds, ndc_poses=ndc_poses, pix_x=pix_x, pix_y=pix_y,
             cam_idx=cam_idx, pixtocam_ndc=pixtocam_ndc,
             dist_keys=np.array(list(dist.keys())), dist_vals=np.array(list(dist.values())))
  cases = {
      'persp': dict(poses=camtoworlds, dist=None, ndc=None, camtype=camera_utils.ProjectionType.PERSPECTIVE),
      'dist': dict(poses=camtoworlds, dist=dist, ndc=None, camtype=camera_utils.ProjectionType.PERSPECTIVE),
      'fisheye': dict(poses=camtoworlds, dist=dist, ndc=None, camtype=camera_utils.ProjectionType.FISHEYE),
      'ndc': dict(poses=ndc_poses, dist=None, ndc=pixtocam_ndc, camtype=camera_utils.ProjectionType.PERSPECTIVE),
      'single': dict(poses=camtoworlds[0], dist=None, ndc=None, camtype=camera_utils.ProjectionType.PERSPECTIVE,
                     p2c=pixtocams[0]),
  }
  meta = lambda v: np.full((B, 1), v, F)
  for name, c in cases.items():
    p2c = c.get('p2c', pixtocams)
    for tag, xnp, dt in [('f64', np, np.float64), ('f32', jnp, np.float32)]:
      cams = (p2c.astype(dt), c['poses'].astype(dt), c['dist'],
              None if c['ndc'] is None else c['ndc'].astype(dt))
      pixels = utils.Pixels(pix_x_int=pix_x, pix_y_int=pix_y, lossmult=meta(1), near=meta(0.2), far=meta(1e6),
                            cam_idx=cam_idx)
      rays = camera_utils.cast_ray_batch(cams, pixels, c['camtype'], xnp=xnp)
      for f in ['origins', 'directions', 'viewdirs', 'radii', 'imageplane']:
        out[f'{name}_{tag}_{f}'] = np.asarray(getattr(rays, f))
  # convert_to_ndc on its own (tests/camera_utils_test.py:27-69 inputs, numpy instead of jax.random)
  o = np.array([0., 0., 1.]) + rng.uniform(-1, 1, (100, 3))
  d = np.array([0., 0., -1.]) + rng.uniform(-.5, .5, (100, 3))
  on, dn = camera_utils.convert_to_ndc(o, d, pixtocam_ndc, 1.0)
  out.update(ndc_in_o=o, ndc_in_d=d, ndc_out_o=on, ndc_out_d=dn)
  path = os.path.join(HERE, 'camera.npz')
  np.savez_compressed(path, **out)
  print('camera.npz', len(out), 'arrays')


if __name__ == '__main__':
  main()
