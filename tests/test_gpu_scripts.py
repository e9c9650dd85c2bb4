"""End to end through the entry scripts (train.py / eval.py / render.py at the repo root) on a tiny
Blender-format scene written to disk: gin bindings -> dataset loader -> training loop with test renders and
summaries -> checkpoint -> evaluation files -> rendered frames.  Needs a B200.

Reference: train.py:43-290, eval.py:44-257, render.py:99-198, internal/datasets.py:507-560."""
import json
import math
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_scene(root, n_train=10, n_test=2, W=40, H=30):
  """A shaded sphere over a gradient sky (train_loop.SyntheticScene.colour), rendered analytically."""
  from PIL import Image
  from multinerf_b200 import camera_utils, train_loop
  angle_x = 0.9
  focal = .5 * W / math.tan(.5 * angle_x)
  p2c = camera_utils.get_pixtocam(focal, W, H)
  for split, n, phase in (('train', n_train, 0.0), ('test', n_test, 0.3)):
    os.makedirs(os.path.join(root, split), exist_ok=True)
    frames = []
    for i in range(n):
      a = 2 * math.pi * (i + phase) / n
      eye = np.array([3.0 * math.cos(a), 3.0 * math.sin(a), 0.5 * math.sin(2 * a)])
      z = eye / np.linalg.norm(eye)
      x = np.cross([0, 0, 1.0], z)
      x /= np.linalg.norm(x)
      y = np.cross(z, x)
      c2w = np.eye(4)
      c2w[:3, :4] = np.concatenate([np.stack([x, y, z], 1), eye[:, None]], 1)
      xs, ys = camera_utils.pixel_coordinates(W, H)
      o, d, v, _, _ = camera_utils.pixels_to_rays(xs, ys, p2c, c2w[:3, :4])
      rgb = train_loop.SyntheticScene.colour(o.reshape(-1, 3), v.reshape(-1, 3)).reshape(H, W, 3).cpu().numpy()
      rgba = np.concatenate([rgb, np.ones((H, W, 1), np.float32)], -1)
      Image.fromarray((rgba * 255 + 0.5).astype(np.uint8)).save(os.path.join(root, split, f'r_{i}.png'))
      frames.append({'file_path': f'./{split}/r_{i}', 'transform_matrix': c2w.tolist()})
    with open(os.path.join(root, f'transforms_{split}.json'), 'w') as f:
      json.dump({'camera_angle_x': angle_x, 'frames': frames}, f)


def _bindings(data_dir, ckpt, steps):
  return [f"Config.data_dir = '{data_dir}'", f"Config.checkpoint_dir = '{ckpt}'", 'Config.batch_size = 1024',
          f'Config.max_steps = {steps}', 'Config.print_every = 20', f'Config.checkpoint_every = {steps}',
          f'Config.train_render_every = {steps // 2}', 'Config.lr_init = 5e-3', 'Config.lr_final = 5e-4',
          'Config.lr_delay_steps = 20', 'Config.render_chunk_size = 512', 'Config.near = 1.5', 'Config.far = 5.0',
          "Config.dataset_loader = 'blender'", 'Model.num_prop_samples = 32', 'Model.num_nerf_samples = 16',
          'PropMLP.net_depth = 2', 'PropMLP.net_width = 64', 'NerfMLP.net_depth = 4', 'NerfMLP.net_width = 128',
          'NerfMLP.bottleneck_width = 64', 'NerfMLP.net_width_viewdirs = 64']


def test_train_eval_render_scripts(tmp_path):
  sys.path.insert(0, ROOT)
  from multinerf_b200 import lib
  lib.require_device()
  import eval as eval_script
  import render as render_script
  import train as train_script
  from multinerf_b200 import checkpoints, train_loop
  data = str(tmp_path / 'scene')
  ckpt = str(tmp_path / 'ckpt')
  _write_scene(data)
  steps = 120
  gin = os.path.join(str(tmp_path), 'mini.gin')
  with open(gin, 'w') as f:            # a config FILE plus bindings, like the reference's command lines
    f.write("PropMLP.disable_density_normals = True\nPropMLP.disable_rgb = True\n"
            "NerfMLP.disable_density_normals = True\nModel.raydist_fn = @jnp.reciprocal\n")
  argv = [f'--gin_configs={gin}'] + [f'--gin_bindings={b}' for b in _bindings(data, ckpt, steps)]
  train_script.main(argv)
  assert checkpoints.latest_checkpoint(ckpt).endswith(f'checkpoint_{steps}')
  assert os.path.exists(os.path.join(ckpt, 'config.gin'))
  log = train_loop.train.summaries.log
  names = {e[1] for e in log}
  for want in ('train_avg_loss', 'train_max_psnr', 'train_learning_rate', 'train_rays_per_sec', 'train_num_params',
               'train_avg_psnr_timed', 'test_rays_per_sec', 'train_metrics/psnr', 'train_metrics/ssim',
               'test_true_color', 'test_output_color', 'test_output_depth_mean', 'test_output_acc'):
    assert want in names, (want, sorted(names))
  test_psnr = [e[2] for e in log if e[1] == 'train_metrics/psnr']
  assert len(test_psnr) == 2 and test_psnr[-1] > 14.0, test_psnr      # the test view is being learned
  # ---- eval.py: the reference's file set under <checkpoint_dir>/test_preds
  eval_script.main(argv)
  out = os.path.join(ckpt, 'test_preds')
  files = set(os.listdir(out))
  for want in ('color_000.png', 'color_cc_000.png', 'color_001.png', 'distance_mean_000.tiff', 'distance_median_001.tiff',
               'acc_000.tiff', f'render_times_{steps}.txt', f'metric_psnr_{steps}.txt', f'metric_ssim_{steps}.txt',
               f'metric_cc_psnr_{steps}.txt'):
    assert want in files, (want, sorted(files))
  psnrs = [float(x) for x in open(os.path.join(out, f'metric_psnr_{steps}.txt')).read().split()]
  assert len(psnrs) == 2 and all(np.isfinite(psnrs)) and abs(psnrs[-1] - test_psnr[-1]) < 6.0
  from multinerf_b200 import utils
  img = utils.load_img(os.path.join(out, 'color_000.png'))
  assert img.shape == (30, 40, 3)
  # ---- render.py
  render_script.main(argv + ['--gin_bindings=Config.render_save_async = False'])
  rdir = os.path.join(ckpt, 'render', f'test_preds_step_{steps}')
  rfiles = set(os.listdir(rdir))
  assert {'color_000.png', 'color_001.png', 'distance_mean_000.tiff', 'distance_median_000.tiff', 'acc_001.tiff'} <= rfiles
  acc = utils.load_img(os.path.join(rdir, 'acc_000.tiff'))
  assert acc.shape == (30, 40) and 0.0 <= float(acc.min()) and float(acc.max()) <= 1.0 + 1e-4


def test_dataset_rays_match_direct_cast(tmp_path):
  """The loader's device-side ray casting equals a direct cast of the same pixels; the test split hands out
  [H, W, n] rays and the ground-truth image."""
  from multinerf_b200 import camera_utils, configs, datasets, lib
  lib.require_device()
  data = str(tmp_path / 'scene')
  _write_scene(data, n_train=3, n_test=1, W=16, H=12)
  cfg = configs.Config(dataset_loader='blender', batch_size=64, near=1.5, far=5.0)
  ds = datasets.load_dataset('train', data, cfg)
  b = next(ds)
  assert b.rays.origins.shape == (64, 1, 1, 3) and b.rays.origins.is_cuda and b.rgb.shape == (64, 1, 1, 3)
  nrm = torch.linalg.norm(b.rays.viewdirs, dim=-1)
  assert float((nrm - 1).abs().max()) < 1e-5 and float(b.rays.near.min()) == 1.5
  dt = datasets.load_dataset('test', data, cfg)
  t = next(dt)
  assert t.rays.origins.shape == (12, 16, 3) and t.rgb.shape == (12, 16, 3)
  xs, ys = camera_utils.pixel_coordinates(16, 12)
  o, d, v, r, ip = camera_utils.pixels_to_rays(xs, ys, dt.pixtocams, dt.camtoworlds[0])
  assert torch.equal(o, t.rays.origins) and torch.equal(d, t.rays.directions) and torch.equal(r, t.rays.radii)
  # cast_rays_in_train_step: the batch carries pixels, the train step casts them
  cfg2 = configs.Config(dataset_loader='blender', batch_size=64, near=1.5, far=5.0, cast_rays_in_train_step=True)
  from multinerf_b200 import utils
  b2 = next(datasets.load_dataset('train', data, cfg2))
  assert isinstance(b2.rays, utils.Pixels)
