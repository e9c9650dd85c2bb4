"""Micro-benchmark of mnrf_pixels_to_rays (run on a B200): time per launch and achieved HBM GB/s
against its algorithmic bytes (12 B read + 48 B written per ray)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinerf_b200 import camera_utils, utils  # noqa: E402


def main():
  rng = np.random.default_rng(0)
  n_cam = 200
  p2c = np.stack([camera_utils.get_pixtocam(f, 1560, 1040) for f in rng.uniform(900, 1100, n_cam)])
  c2w = np.concatenate([np.broadcast_to(np.eye(3), (n_cam, 3, 3)), rng.uniform(-1, 1, (n_cam, 3, 1))], -1)
  dist = dict(k1=0.05, k2=-0.02, k3=0.004, k4=0.0, p1=0.001, p2=-0.0015)
  for B in [16384, 1 << 20, 1 << 24]:
    px = torch.tensor(rng.integers(0, 1560, B).astype(np.int32)).cuda()
    py = torch.tensor(rng.integers(0, 1040, B).astype(np.int32)).cuda()
    cam = torch.tensor(rng.integers(0, n_cam, (B, 1)).astype(np.int32)).cuda()
    meta = torch.ones(B, 1, device='cuda')
    pixels = utils.Pixels(px, py, meta, meta, meta, cam)
    for name, d in [('pinhole', None), ('undistort', dist)]:
      cams = (torch.tensor(p2c, dtype=torch.float32).cuda(), torch.tensor(c2w, dtype=torch.float32).cuda(), d, None)
      for _ in range(3):
        camera_utils.cast_ray_batch(cams, pixels)
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      iters = 20
      e0.record()
      for _ in range(iters):
        camera_utils.cast_ray_batch(cams, pixels)
      e1.record()
      torch.cuda.synchronize()
      ms = e0.elapsed_time(e1) / iters
      print(f'rays={B:9d} {name:10s} {ms * 1e3:9.1f} us/launch (incl. 5 output allocations)  '
            f'{B / ms / 1e6:8.2f} G rays/s  {60.0 * B / ms / 1e6:8.1f} GB/s algorithmic', flush=True)


if __name__ == '__main__':
  main()
