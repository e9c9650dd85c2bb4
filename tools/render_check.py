"""Multi-GPU check of the render path: every rank renders its slice, pixels are all-gathered
(NCCL), and rank 0 compares against a single-GPU render of the same image."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinerf_b200 import configs, models, train_utils, utils  # noqa: E402


def main():
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  H, W = 96, 131                       # 12576 rays: one 8192 chunk + a ragged one (needs padding)
  bundle = configs.bundle_360()
  bundle.config.render_chunk_size = 8192
  f = np.float32
  ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
  d = np.stack([(xs - W / 2) / 120.0, (ys - H / 2) / 120.0, -np.ones_like(xs, dtype=np.float64)], -1)
  v = d / np.linalg.norm(d, axis=-1, keepdims=True)
  o = np.broadcast_to(np.array([0.5, 0.5, 0.3]), d.shape)
  rays = utils.Rays(origins=o.astype(f), directions=d.astype(f), viewdirs=v.astype(f),
                    radii=np.full((H, W, 1), 7e-4, f), imageplane=np.zeros((H, W, 2), f),
                    lossmult=np.ones((H, W, 1), f), near=np.full((H, W, 1), 0.2, f),
                    far=np.full((H, W, 1), 1e6, f), cam_idx=np.zeros((H, W, 1), np.int32))
  model, state, render_eval_pfn, train_pstep, lr_fn = train_utils.setup_model(bundle, 0)
  render_fn = lambda rng, r: render_eval_pfn(state.params, 1.0, None, r)
  out = models.render_image(render_fn, rays, None, bundle, verbose=False, world_size=world, rank=rank)
  torch.cuda.synchronize()
  if rank == 0:
    assert out['rgb'].shape == (H, W, 3), out['rgb'].shape
    # single-GPU reference on rank 0 (bypass the gather)
    single = models.render_image(lambda rng, r: model.apply(state.params, None, r, 1.0, True), rays, None, bundle,
                                 verbose=False)
    for k in ['rgb', 'acc', 'distance_mean', 'distance_median']:
      err = float((out[k] - single[k]).abs().max())
      print(f'render {world} GPU vs 1 GPU: max |d {k}| = {err:.3e}')
      assert err < 1e-5, (k, err)
    print('render check ok', tuple(out['rgb'].shape))
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
