"""BASELINE config 5: render a 1560x1040 image with the 360.gin model, rays sharded over the
ranks, pixels all-gathered per 16384-ray chunk (NCCL).  Prints one JSON line on rank 0.

  python tools/render_bench.py                      # 1 GPU
  torchrun --nproc-per-node 8 tools/render_bench.py # 8 GPUs
MLP queries per ray actually run: 64 + 64 proposal + 32 NeRF = 160 (no reference config evaluates
"128 samples/ray").
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinerf_b200 import configs, models, train_utils, utils  # noqa: E402

FWD_FLOP_PER_RAY = 638435328


def main():
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  W, H = 1560, 1040
  bundle = configs.bundle_360()
  f = np.float32
  ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing='ij')
  cam = np.array([0.5, 0.5, 0.3])
  fwd = -cam / np.linalg.norm(cam)
  right = np.cross(fwd, [0, 0, 1.0]); right /= np.linalg.norm(right)
  up = np.cross(right, fwd)
  d = ((xs - W / 2)[..., None] * right + (H / 2 - ys)[..., None] * up) / 1200.0 + fwd
  v = d / np.linalg.norm(d, axis=-1, keepdims=True)
  rays = utils.Rays(origins=np.broadcast_to(cam, d.shape).astype(f), directions=d.astype(f),
                    viewdirs=v.astype(f), radii=np.full((H, W, 1), 2 / (1200 * np.sqrt(12)), f),
                    imageplane=np.zeros((H, W, 2), f), lossmult=np.ones((H, W, 1), f),
                    near=np.full((H, W, 1), 0.2, f), far=np.full((H, W, 1), 1e6, f),
                    cam_idx=np.zeros((H, W, 1), np.int32))
  # rays resident on the device (the reference feeds host rays per chunk; both are timed)
  dev_rays = rays.map(lambda a: torch.as_tensor(a).cuda())
  model, state, render_eval_pfn, _, _ = train_utils.setup_model(bundle, 0)
  render_fn = lambda rng, r: render_eval_pfn(state.params, 1.0, None, r)

  def run(r):
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    t0 = time.perf_counter()
    out = models.render_image(render_fn, r, None, bundle, verbose=False, world_size=world, rank=rank)
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    return time.perf_counter() - t0, out
  run(dev_rays)                       # warm-up (allocations, tensor maps)
  t_dev, out = run(dev_rays)
  t_host, _ = run(rays)
  if rank == 0:
    n = H * W
    print(json.dumps({
        'metric': 'render 1560x1040 (360.gin, 160 MLP queries/ray), seconds per image', 'n_gpus': world,
        'value': t_dev, 'unit': 's/image', 'rays_per_s': n / t_dev, 'higher_is_better': False,
        'host_rays_s_per_image': t_host,
        'fwd_tflops_per_gpu': n * FWD_FLOP_PER_RAY / t_dev / world / 1e12,
        'outputs': sorted(k for k in out if not k.startswith('ray_')), 'rgb_shape': list(out['rgb'].shape)}))
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
