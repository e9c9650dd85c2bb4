#!/usr/bin/env python
"""Training entry point (counterpart of the reference's train.py).

  python train.py --gin_configs=configs/blender_256.gin \
      --gin_bindings="Config.data_dir = '/data/nerf_synthetic/lego'" \
      --gin_bindings="Config.checkpoint_dir = '/tmp/lego'"
  torchrun --nproc-per-node 8 train.py ...          (one process per GPU, NCCL)
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from multinerf_b200 import configs, datasets, train_loop  # noqa: E402


def parse(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument('--gin_configs', action='append', default=[])
  ap.add_argument('--gin_bindings', action='append', default=[])
  ap.add_argument('--no_graph', action='store_true', help='launch kernels eagerly instead of replaying a CUDA graph')
  return ap.parse_args(argv)


def setup_distributed():
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  if world > 1 and not dist.is_initialized():
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  return world, rank, torch.device('cuda', local)


def main(argv=None):
  args = parse(argv)
  world, rank, device = setup_distributed()
  bundle = configs.load_config(args.gin_configs, args.gin_bindings, search_paths=[ROOT, os.getcwd()])
  config = bundle.config
  if config.batch_size % world != 0:
    raise ValueError('Batch size must be divisible by the number of devices.')
  np.random.seed(20201473 + rank)
  if config.checkpoint_dir and rank == 0:
    os.makedirs(config.checkpoint_dir, exist_ok=True)
    with open(os.path.join(config.checkpoint_dir, 'config.gin'), 'w') as f:      # configs.py:149-152
      for path in args.gin_configs:
        f.write(f"include '{path}'\n")
      f.write('\n'.join(args.gin_bindings) + '\n')
  dataset = datasets.load_dataset('train', config.data_dir, config, device=device, rank=rank, world=world)
  test_dataset = datasets.load_dataset('test', config.data_dir, config, device=device, rank=rank, world=world)
  train_loop.train(bundle, dataset, use_graph=not args.no_graph, test_dataset=test_dataset)
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
