"""Train a config on the procedural scene of multinerf_b200.train_loop.SyntheticScene and print the
reference's summary lines (run on a B200).

  python tools/train_synthetic.py [--steps 300] [--batch 16384] [--mini] [--graph] [--cast_rays] [--ckpt DIR]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinerf_b200 import configs, train_loop  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--steps', type=int, default=300)
  ap.add_argument('--batch', type=int, default=16384)
  ap.add_argument('--mini', action='store_true', help='PropMLP 2x64 / NerfMLP 6x128 instead of 360.gin sizes')
  ap.add_argument('--graph', action='store_true')
  ap.add_argument('--cast_rays', action='store_true', help='Config.cast_rays_in_train_step: feed pixels + cameras')
  ap.add_argument('--ckpt', default=None)
  args = ap.parse_args()
  b = configs.bundle_360()
  if args.mini:
    b.model.num_prop_samples, b.model.num_nerf_samples = 32, 16
    b.prop_mlp.net_depth, b.prop_mlp.net_width = 2, 64
    b.nerf_mlp.net_depth, b.nerf_mlp.net_width = 6, 128
    b.nerf_mlp.bottleneck_width, b.nerf_mlp.net_width_viewdirs = 64, 64
  c = b.config
  c.batch_size, c.max_steps, c.print_every = args.batch, args.steps, max(1, args.steps // 10)
  c.lr_delay_steps = min(c.lr_delay_steps, args.steps // 4)
  c.checkpoint_every, c.checkpoint_dir = max(1, args.steps // 2), args.ckpt
  c.cast_rays_in_train_step = args.cast_rays
  ds = train_loop.SyntheticScene(c)
  train_loop.train(b, ds, use_graph=args.graph)


if __name__ == '__main__':
  main()
