"""Pixel -> ray generation on the device (reference surface: internal/camera_utils.py).

`cast_ray_batch(cameras, pixels, camtype)` and `pixels_to_rays(...)` keep the reference's names,
argument meaning and return order (camera_utils.py:522-688); the work is one launch of
`mnrf_pixels_to_rays` (csrc/camera.cu).  The small host helpers (`intrinsic_matrix`,
`get_pixtocam`, `pixel_coordinates`) are the reference's one-liners in numpy.
"""
import enum

import numpy as np
import torch

from . import lib as L
from . import utils


class ProjectionType(enum.Enum):
  """camera_utils.py:516-519."""
  PERSPECTIVE = 'perspective'
  FISHEYE = 'fisheye'


def intrinsic_matrix(fx, fy, cx, cy):
  """camera_utils.py:398-408."""
  return np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.]])


def get_pixtocam(focal, width, height):
  """camera_utils.py:411-417."""
  return np.linalg.inv(intrinsic_matrix(focal, focal, width * .5, height * .5))


def pixel_coordinates(width, height):
  """camera_utils.py:420-424."""
  return np.meshgrid(np.arange(width), np.arange(height), indexing='xy')


def _dev(x, dtype, device):
  t = x if isinstance(x, torch.Tensor) else torch.tensor(np.asarray(x))
  return t.to(device=device, dtype=dtype, non_blocking=True).contiguous()


_checked = False


def _launch(pix_x, pix_y, cam_idx, pixtocams, camtoworlds, distortion_params, pixtocam_ndc, camtype):
  """All inputs flat on the device; returns the five [B, n] fp32 outputs."""
  global _checked
  if not _checked:
    L.require_device()          # queries device properties: once, not per launch
    _checked = True
  lib = L.load()
  B = pix_x.shape[0]
  dev = pix_x.device
  if isinstance(camtype, str):
    camtype = ProjectionType(camtype)
  if camtype not in (ProjectionType.PERSPECTIVE, ProjectionType.FISHEYE):
    raise ValueError(f'unknown camtype {camtype!r}')
  dp = dict(distortion_params or {})
  unknown = set(dp) - {'k1', 'k2', 'k3', 'k4', 'p1', 'p2'}
  if unknown:
    raise TypeError(f'unexpected distortion parameters {sorted(unknown)}')   # as **distortion_params would
  ndc = None if pixtocam_ndc is None else np.asarray(
      pixtocam_ndc.detach().cpu() if isinstance(pixtocam_ndc, torch.Tensor) else pixtocam_ndc, np.float64)
  d = L.CameraDesc(B, pixtocams.shape[0], 0 if camtype == ProjectionType.PERSPECTIVE else 1,
                   int(distortion_params is not None),
                   float(dp.get('k1', 0.0)), float(dp.get('k2', 0.0)), float(dp.get('k3', 0.0)),
                   float(dp.get('k4', 0.0)), float(dp.get('p1', 0.0)), float(dp.get('p2', 0.0)),
                   1e-9, 10, int(ndc is not None),
                   float(ndc[0, 2]) if ndc is not None else 1.0, float(ndc[1, 2]) if ndc is not None else 1.0,
                   1.0)
  out = [torch.empty(B, n, device=dev, dtype=torch.float32) for n in (3, 3, 3, 1, 2)]
  from . import ops
  ops._count()
  L.check(lib.mnrf_pixels_to_rays(L.C.byref(d), L.ptr(pix_x), L.ptr(pix_y), L.ptr(cam_idx), L.ptr(pixtocams),
                                  L.ptr(camtoworlds), *[L.ptr(t) for t in out], L.stream_ptr()))
  return out


def pixels_to_rays(pix_x_int, pix_y_int, pixtocams, camtoworlds, distortion_params=None,
                   pixtocam_ndc=None, camtype=ProjectionType.PERSPECTIVE, device='cuda'):
  """camera_utils.py:522-636.  `pixtocams` / `camtoworlds` are one matrix ([3,3] / [3,4]) or one per
  pixel (SH + [3,3] / SH + [3,4]); returns (origins, directions, viewdirs, radii, imageplane) with
  shapes SH + [3|3|3|1|2] as CUDA fp32 tensors."""
  px = _dev(pix_x_int, torch.int32, device)
  sh = tuple(px.shape)
  px = px.reshape(-1)
  py = _dev(pix_y_int, torch.int32, device).reshape(-1)
  p2c = _dev(pixtocams, torch.float32, device)
  c2w = _dev(camtoworlds, torch.float32, device)
  B = px.shape[0]
  if p2c.ndim == 2 and c2w.ndim == 2:
    p2c, c2w, idx = p2c[None], c2w[None, :3, :4].contiguous(), None
  else:
    p2c = p2c.expand(sh + (3, 3)).reshape(B, 3, 3).contiguous()
    c2w = c2w[..., :3, :4].expand(sh + (3, 4)).reshape(B, 3, 4).contiguous()
    idx = torch.arange(B, device=px.device, dtype=torch.int32)
  outs = _launch(px, py, idx, p2c, c2w, distortion_params, pixtocam_ndc, camtype)
  return tuple(o.reshape(sh + (o.shape[-1],)) for o in outs)


def cast_ray_batch(cameras, pixels, camtype=ProjectionType.PERSPECTIVE, device='cuda'):
  """camera_utils.py:639-688: cameras = (pixtocams, camtoworlds, distortion_params, pixtocam_ndc) with
  1 or N stacked matrices; pixels = utils.Pixels.  The per-ray camera gather happens in the kernel."""
  pixtocams, camtoworlds, distortion_params, pixtocam_ndc = cameras
  px = _dev(pixels.pix_x_int, torch.int32, device)
  sh = tuple(px.shape)
  px = px.reshape(-1)
  py = _dev(pixels.pix_y_int, torch.int32, device).reshape(-1)
  idx = _dev(pixels.cam_idx, torch.int32, device).reshape(-1)
  p2c = _dev(pixtocams, torch.float32, device)
  c2w = _dev(camtoworlds, torch.float32, device)
  p2c = p2c[None] if p2c.ndim == 2 else p2c.reshape(-1, 3, 3)
  c2w = (c2w[None] if c2w.ndim == 2 else c2w.reshape((-1,) + tuple(c2w.shape[-2:])))[:, :3, :4].contiguous()
  n_cam = max(p2c.shape[0], c2w.shape[0])
  if p2c.shape[0] != n_cam:
    p2c = p2c.expand(n_cam, 3, 3).contiguous()
  if c2w.shape[0] != n_cam:
    c2w = c2w.expand(n_cam, 3, 4).contiguous()
  o, d, v, r, ip = _launch(px, py, idx if n_cam > 1 else None, p2c, c2w, distortion_params, pixtocam_ndc, camtype)
  rs = lambda t: t.reshape(sh + (t.shape[-1],))
  return utils.Rays(origins=rs(o), directions=rs(d), viewdirs=rs(v), radii=rs(r), imageplane=rs(ip),
                    lossmult=pixels.lossmult, near=pixels.near, far=pixels.far, cam_idx=pixels.cam_idx,
                    exposure_idx=pixels.exposure_idx, exposure_values=pixels.exposure_values)
