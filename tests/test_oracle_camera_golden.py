"""Oracle pixel->ray generation vs the reference's own internal/camera_utils.py outputs
(tests/golden/camera.npz, produced by tests/golden/make_golden_camera.py) and the reference's
known-answer test tests/camera_utils_test.py:27-69."""
import math
import os
import types

import numpy as np
import pytest
import torch

from oracle import o_camera

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'camera.npz'))
FIELDS = ['origins', 'directions', 'viewdirs', 'radii', 'imageplane']


def _case(name, dtype):
  dist = None
  if name in ('dist', 'fisheye'):
    dist = {str(k): float(v) for k, v in zip(G['dist_keys'], G['dist_vals'])}
  poses = G['ndc_poses'] if name == 'ndc' else G['camtoworlds']
  p2c = G['pixtocams']
  if name == 'single':
    poses, p2c = poses[0], p2c[0]
  ndc = torch.tensor(G['pixtocam_ndc'], dtype=dtype) if name == 'ndc' else None
  cams = (torch.tensor(p2c, dtype=dtype), torch.tensor(poses, dtype=dtype), dist, ndc)
  pixels = types.SimpleNamespace(pix_x_int=torch.tensor(G['pix_x']), pix_y_int=torch.tensor(G['pix_y']),
                                 cam_idx=torch.tensor(G['cam_idx']))
  camtype = o_camera.FISHEYE if name == 'fisheye' else o_camera.PERSPECTIVE
  return cams, pixels, camtype


@pytest.mark.parametrize('name', ['persp', 'dist', 'fisheye', 'ndc', 'single'])
def test_cast_ray_batch_vs_reference(name):
  for tag, dtype, tol in [('f64', torch.float64, 1e-12), ('f32', torch.float32, 2e-5)]:
    cams, pixels, camtype = _case(name, dtype)
    out = o_camera.cast_ray_batch(cams, pixels, camtype)
    for f in FIELDS:
      ref = G[f'{name}_{tag}_{f}']
      got = out[f].numpy()
      assert got.shape == ref.shape, (name, tag, f, got.shape, ref.shape)
      scale = max(1.0, float(np.abs(ref).max()))
      err = float(np.abs(got.astype(np.float64) - ref.astype(np.float64)).max())
      # radii in fp32 come from differences of nearly equal directions: relative tolerance
      lim = tol * scale if f != 'radii' else max(tol * scale, 2e-3 * float(np.abs(ref).max()))
      assert err <= lim, (name, tag, f, err, lim)
  # the fp32 path agrees with the fp64 answer to fp32 accuracy (what the device kernel is held to)
  cams, pixels, camtype = _case(name, torch.float32)
  out = o_camera.cast_ray_batch(cams, pixels, camtype)
  for f in ['origins', 'directions', 'viewdirs', 'imageplane']:
    ref = G[f'{name}_f64_{f}']
    assert np.abs(out[f].numpy() - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), (name, f)


def test_convert_to_ndc_golden_and_known_answer():
  o, d = torch.tensor(G['ndc_in_o']), torch.tensor(G['ndc_in_d'])
  on, dn = o_camera.convert_to_ndc(o, d, torch.tensor(G['pixtocam_ndc']), 1.0)
  np.testing.assert_allclose(on.numpy(), G['ndc_out_o'], atol=1e-12)
  np.testing.assert_allclose(dn.numpy(), G['ndc_out_d'], atol=1e-12)
  # tests/camera_utils_test.py:27-69: world points along each ray project onto the NDC ray
  rng = np.random.default_rng(0)
  for _ in range(10):
    focal, width, height = rng.uniform(100.0, 200.0, 3)
    pixtocam = o_camera.get_pixtocam(focal, width, height)
    near = 1.0
    origins = torch.tensor([0.0, 0.0, 1.0]) + torch.tensor(rng.uniform(-1, 1, (1000, 3)))
    directions = torch.tensor([0.0, 0.0, -1.0]) + torch.tensor(rng.uniform(-0.5, 0.5, (1000, 3)))
    t = torch.linspace(0.0, 1.0, 10, dtype=torch.float64)
    pts_world = origins + t[:, None, None] * directions
    pts_ndc = torch.stack([-focal / (0.5 * width) * pts_world[..., 0] / pts_world[..., 2],
                           -focal / (0.5 * height) * pts_world[..., 1] / pts_world[..., 2],
                           1.0 + 2.0 * near / pts_world[..., 2]], dim=-1)
    o_ndc, d_ndc = o_camera.convert_to_ndc(origins, directions, pixtocam, near)
    unit = d_ndc / torch.linalg.norm(d_ndc, dim=-1, keepdim=True)
    proj = ((pts_ndc - o_ndc) * unit).sum(dim=-1)
    np.testing.assert_allclose(pts_ndc.numpy(), (o_ndc + unit * proj[..., None]).numpy(), atol=1e-5, rtol=1e-5)


def test_undistort_inverts_distortion():
  rng = np.random.default_rng(3)
  x = torch.tensor(rng.uniform(-0.6, 0.6, 500))
  y = torch.tensor(rng.uniform(-0.6, 0.6, 500))
  k = dict(k1=0.05, k2=-0.02, k3=0.004, k4=0.0, p1=0.001, p2=-0.0015)
  fx, fy, *_ = o_camera._residual_and_jacobian(x, y, torch.zeros_like(x), torch.zeros_like(y), **k)
  xu, yu = o_camera.radial_and_tangential_undistort(fx, fy, **k)      # fx, fy = distorted coordinates
  np.testing.assert_allclose(xu.numpy(), x.numpy(), atol=1e-9)
  np.testing.assert_allclose(yu.numpy(), y.numpy(), atol=1e-9)
