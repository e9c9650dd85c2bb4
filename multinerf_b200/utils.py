"""Input/output containers of the hot path (reference: internal/utils.py:31-136).

`Rays` / `Batch` keep the reference's field names.  Arrays may be numpy or torch, with
arbitrary leading dims (`[B, n]`, `[H, W, n]`, `[B, 1, 1, n]`, ...); the device path flattens
them to `[B, n]` fp32.
"""
import dataclasses
from typing import Any, Optional

import numpy as np
import torch


@dataclasses.dataclass
class Pixels:
  """internal/utils.py:31-41: integer pixel coordinates [SH] + per-ray metadata [SH, 1]."""
  pix_x_int: Any
  pix_y_int: Any
  lossmult: Any
  near: Any
  far: Any
  cam_idx: Any
  exposure_idx: Optional[Any] = None
  exposure_values: Optional[Any] = None


@dataclasses.dataclass
class Rays:
  origins: Any
  directions: Any
  viewdirs: Any
  radii: Any
  imageplane: Any
  lossmult: Any
  near: Any
  far: Any
  cam_idx: Any
  exposure_idx: Optional[Any] = None
  exposure_values: Optional[Any] = None

  def map(self, fn):
    kw = {}
    for f in dataclasses.fields(self):
      v = getattr(self, f.name)
      kw[f.name] = None if v is None else fn(v)
    return Rays(**kw)


@dataclasses.dataclass
class Batch:
  rays: Rays
  rgb: Optional[Any] = None
  disps: Optional[Any] = None
  normals: Optional[Any] = None
  alphas: Optional[Any] = None


def dummy_rays(include_exposure_idx=False, include_exposure_values=False):
  """internal/utils.py:60-79."""
  z = lambda n: np.zeros((1, n), np.float32)
  kw = {}
  if include_exposure_idx:
    kw['exposure_idx'] = z(1).astype(np.int32)
  if include_exposure_values:
    kw['exposure_values'] = z(1)
  return Rays(origins=z(3), directions=z(3), viewdirs=z(3), radii=z(1), imageplane=z(2),
              lossmult=z(1), near=z(1), far=z(1), cam_idx=z(1).astype(np.int32), **kw)


def to_device_flat(rays, device):
  """Flatten leading dims to [B, n], move to `device` as contiguous fp32 (int32 for indices)."""
  def conv(name, v):
    t = torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v)
    t = t.reshape(-1, t.shape[-1])
    if name in ('cam_idx', 'exposure_idx'):
      return t.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
    return t.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
  kw = {}
  for f in dataclasses.fields(rays):
    v = getattr(rays, f.name)
    kw[f.name] = None if v is None else conv(f.name, v)
  return Rays(**kw)


def shard(xs, num_shards):
  """Split the leading dim into [num_shards, -1, ...] (internal/utils.py:125-128)."""
  def one(x):
    return x.reshape((num_shards, -1) + tuple(x.shape[1:]))
  if isinstance(xs, Rays):
    return xs.map(one)
  if isinstance(xs, dict):
    return {k: shard(v, num_shards) for k, v in xs.items()}
  return one(xs)


def unshard(x, padding=0):
  """internal/utils.py:131-136."""
  y = x.reshape((x.shape[0] * x.shape[1],) + tuple(x.shape[2:]))
  if padding > 0:
    y = y[:-padding]
  return y


# ---------------------------------------------------------------------------------------------------
# File / image IO used by the dataset loaders and the eval / render scripts (internal/utils.py:90-171)
# ---------------------------------------------------------------------------------------------------
import enum  # noqa: E402
import os  # noqa: E402


class DataSplit(enum.Enum):
  TRAIN = 'train'
  TEST = 'test'


class BatchingMethod(enum.Enum):
  """Rays of a batch come from all images or from one image (internal/utils.py:96-99)."""
  ALL_IMAGES = 'all_images'
  SINGLE_IMAGE = 'single_image'


def open_file(pth, mode='r'):
  return open(pth, mode=mode)


def file_exists(pth):
  return os.path.exists(pth)


def listdir(pth):
  return os.listdir(pth)


def isdir(pth):
  return os.path.isdir(pth)


def makedirs(pth):
  os.makedirs(pth, exist_ok=True)


def load_img(pth):
  """Image file -> float32 array with the file's own value range (internal/utils.py:134-138)."""
  from PIL import Image
  with open(pth, 'rb') as f:
    return np.array(Image.open(f), dtype=np.float32)


def load_exif(pth):
  """EXIF tags by name, {} when the file has none (internal/utils.py:141-153)."""
  from PIL import ExifTags, Image
  with open(pth, 'rb') as f:
    raw = Image.open(f)._getexif()  # pylint: disable=protected-access
  return {} if raw is None else {ExifTags.TAGS[k]: v for k, v in raw.items() if k in ExifTags.TAGS}


def _to_numpy(x):
  return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def save_img_u8(img, pth):
  """[0, 1] image -> 8-bit PNG (NaN -> 0, clipped) (internal/utils.py:156-162)."""
  from PIL import Image
  arr = (np.clip(np.nan_to_num(_to_numpy(img)), 0., 1.) * 255.).astype(np.uint8)
  with open(pth, 'wb') as f:
    Image.fromarray(arr).save(f, 'PNG')


def save_img_f32(depthmap, pth):
  """Float map (distance, acc) -> float32 TIFF (internal/utils.py:165-168)."""
  from PIL import Image
  with open(pth, 'wb') as f:
    Image.fromarray(np.nan_to_num(_to_numpy(depthmap)).astype(np.float32)).save(f, 'TIFF')
