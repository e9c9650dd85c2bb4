"""The reference's training loop around `train_pstep` (train.py:66-223), for any iterable of
utils.Batch: same step/`train_frac` bookkeeping, the same summary line every `print_every` steps
(`loss`, `psnr`, `lr`, each `losses/*` term, `r/s` = batch_size * steps/s), checkpoints at step 1,
every `checkpoint_every` and on exit, resume from the newest checkpoint.

`SyntheticScene` is a procedural dataset (a shaded sphere over a gradient background seen from
cameras on a ring) used by the tests and `tools/train_synthetic.py`: the reference's loaders need
scene files that are not in this container.
"""
import gc
import math
import time

import numpy as np
import torch

from . import camera_utils
from . import checkpoints
from . import train_utils
from . import utils


class SyntheticScene:
  """Infinite iterator of utils.Batch for a procedurally defined scene.

  rays: utils.Pixels when `cast_rays_in_train_step` (the cameras are in `.cameras`), else utils.Rays
  generated on the device by camera_utils.cast_ray_batch; rgb: [B, 3] analytic colours."""

  def __init__(self, config, n_cameras=24, width=96, height=72, focal=90.0, radius=3.0, seed=0,
               device='cuda', rank=0, world=1):
    self.config = config
    self.width, self.height = width, height
    self.rng = np.random.default_rng(seed + 7919 * rank)
    self.batch = config.batch_size // world
    self.device = device
    self.camtype = camera_utils.ProjectionType.PERSPECTIVE
    self.size = n_cameras
    self.metadata = None
    p2c = camera_utils.get_pixtocam(focal, width, height)
    poses = []
    for i in range(n_cameras):
      a = 2 * math.pi * i / n_cameras
      eye = np.array([radius * math.cos(a), radius * math.sin(a), 0.6 * math.sin(2 * a)])
      z = eye / np.linalg.norm(eye)                      # camera looks down -z at the origin
      x = np.cross(np.array([0.0, 0.0, 1.0]), z)
      x /= np.linalg.norm(x)
      y = np.cross(z, x)
      poses.append(np.concatenate([np.stack([x, y, z], 1), eye[:, None]], 1))
    self.cameras = (np.broadcast_to(p2c, (n_cameras, 3, 3)).copy(), np.stack(poses), None, None)
    self._dev_cameras = (torch.tensor(self.cameras[0], dtype=torch.float32, device=device),
                         torch.tensor(self.cameras[1], dtype=torch.float32, device=device), None, None)

  @staticmethod
  def colour(origins, viewdirs):
    """Analytic radiance: unit-free shaded sphere of radius 0.8 at the origin, else a gradient sky."""
    o, d = origins, viewdirs
    b = (o * d).sum(-1)
    c = (o * o).sum(-1) - 0.64
    disc = b * b - c
    hit = disc > 0
    t = -b - torch.sqrt(disc.clamp(min=0))
    n = torch.nn.functional.normalize(o + t[:, None] * d, dim=-1)
    sphere = 0.5 + 0.5 * n * torch.tensor([1.0, 0.8, 0.6], device=o.device)
    sky = torch.stack([0.15 + 0.1 * d[:, 2], 0.2 + 0.15 * d[:, 2], 0.45 + 0.3 * d[:, 2]], -1)
    return torch.where((hit & (t > 0))[:, None], sphere, sky).clamp(0, 1)

  def pixels(self):
    B = self.batch
    meta = lambda v: np.full((B, 1), v, np.float32)
    return utils.Pixels(pix_x_int=self.rng.integers(0, self.width, B).astype(np.int32),
                        pix_y_int=self.rng.integers(0, self.height, B).astype(np.int32),
                        lossmult=meta(1.0), near=meta(self.config.near), far=meta(self.config.far),
                        cam_idx=self.rng.integers(0, self.size, (B, 1)).astype(np.int32))

  def __iter__(self):
    return self

  def __next__(self):
    px = self.pixels()
    rays = camera_utils.cast_ray_batch(self._dev_cameras, px, self.camtype, device=self.device)
    rgb = self.colour(rays.origins, rays.viewdirs)
    return utils.Batch(rays=px if self.config.cast_rays_in_train_step else rays, rgb=rgb)


def train(bundle, dataset, seed=20200823, log=print, use_graph=False):
  """train.py:66-223 without TensorBoard / test-set rendering.  Returns (model, state, history) where
  history is the list of summary dicts printed every `print_every` steps."""
  config = bundle.config
  model, state, _, _, lr_fn = train_utils.setup_model(bundle, seed, dataset=dataset)
  train_pstep = train_utils.create_train_step(model, config, use_graph=use_graph, dataset=dataset)
  world, rank = train_utils._world()
  if rank == 0:
    log(f'Number of parameters being optimized: {model.num_params()}')
  if getattr(dataset, 'size', 0) > model.num_glo_embeddings and model.num_glo_features > 0:
    raise ValueError(f'Number of glo embeddings {model.num_glo_embeddings} must be at least equal to '
                     f'number of train images {dataset.size}')                      # train.py:74-78
  if config.checkpoint_dir:
    state = checkpoints.restore_checkpoint(config.checkpoint_dir, state, model=model)
  init_step = state.step + 1
  cameras = getattr(dataset, 'cameras', None)
  gen = torch.Generator(device=model.device)
  gen.manual_seed(seed + rank)                       # separate random streams per process (train.py:103)
  num_steps = config.early_exit_steps if config.early_exit_steps is not None else config.max_steps
  history, stats_buffer = [], []
  reset_stats, train_start = True, time.time()
  gc.disable()
  try:
    for step, batch in zip(range(init_step, num_steps + 1), dataset):
      if reset_stats:
        stats_buffer, train_start, reset_stats = [], time.time(), False
      learning_rate = lr_fn(step)
      train_frac = float(np.clip((step - 1) / max(1, config.max_steps - 1), 0, 1))
      state, stats, gen = train_pstep(gen, state, batch, cameras, train_frac, 1.0)
      stats_buffer.append(stats)
      if step % 10000 == 0:
        gc.collect()
      if step == init_step or step % config.print_every == 0:
        torch.cuda.synchronize()
        elapsed = time.time() - train_start
        steps_per_sec = len(stats_buffer) / elapsed
        rays_per_sec = config.batch_size * steps_per_sec
        mats = [s.materialize() for s in stats_buffer]
        avg = {'loss': float(np.mean([m['loss'] for m in mats])), 'psnr': float(np.mean([m['psnr'] for m in mats]))}
        enabled = {'interlevel': config.interlevel_loss_mult > 0, 'distortion': config.distortion_loss_mult > 0,
                   'orientation': config.orientation_coarse_loss_mult > 0 or config.orientation_loss_mult > 0,
                   'predicted_normals': (config.predicted_normal_coarse_loss_mult > 0 or
                                         config.predicted_normal_loss_mult > 0)}      # train_utils.py:283-303
        for k in mats[0]['losses']:
          if enabled.get(k, True):
            avg['losses/' + k] = float(np.mean([m['losses'][k] for m in mats]))
        summary = dict(step=step, lr=learning_rate, steps_per_sec=steps_per_sec, rays_per_sec=rays_per_sec, **avg)
        history.append(summary)
        if rank == 0:
          precision = int(np.ceil(np.log10(config.max_steps))) + 1
          str_losses = {k[7:11]: (f'{v:0.5f}' if 1e-4 <= v < 10 else f'{v:0.1e}')
                        for k, v in avg.items() if k.startswith('losses/')}
          log(f'{step:{precision}d}/{config.max_steps:d}: loss={avg["loss"]:0.5f}, psnr={avg["psnr"]:6.3f}, '
              f'lr={learning_rate:0.2e} | ' + ', '.join(f'{k}={s}' for k, s in str_losses.items()) +
              f', {rays_per_sec:0.0f} r/s')
        reset_stats = True
      if config.checkpoint_dir and rank == 0 and (step == 1 or step % config.checkpoint_every == 0):
        checkpoints.save_checkpoint(config.checkpoint_dir, state, int(step), keep=100, model=model)
    if config.checkpoint_dir and rank == 0 and config.max_steps % config.checkpoint_every != 0:
      checkpoints.save_checkpoint(config.checkpoint_dir, state, int(config.max_steps), keep=100, model=model)   # train.py:284-287
  finally:
    gc.enable()
  return model, state, history
