#!/usr/bin/env python
"""Evaluation entry point (counterpart of the reference's eval.py): renders the test set from the newest
checkpoint and writes images + metric files under <checkpoint_dir>/test_preds.

  python eval.py --gin_configs=configs/blender_256.gin --gin_bindings="Config.data_dir = '...'" \
      --gin_bindings="Config.checkpoint_dir = '...'"
With `Config.eval_only_once = False` it keeps polling the checkpoint directory like the reference.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from multinerf_b200 import checkpoints, configs, datasets, eval_lib, train_loop  # noqa: E402
from train import parse, setup_distributed  # noqa: E402


def main(argv=None):
  args = parse(argv)
  world, rank, device = setup_distributed()
  bundle = configs.load_config(args.gin_configs, args.gin_bindings, search_paths=[ROOT, os.getcwd()])
  config = bundle.config
  dataset = datasets.load_dataset('test', config.data_dir, config, device=device, rank=rank, world=world)
  summaries = None
  if not config.eval_only_once and rank == 0:
    summaries = train_loop._Summaries(os.path.join(config.checkpoint_dir, 'eval'))
  last_step = 0
  while True:
    latest = checkpoints._steps(config.checkpoint_dir)
    step = latest[-1] if latest else 0
    if step <= last_step:
      print(f'Checkpoint step {step} <= last step {last_step}, sleeping.')
      time.sleep(10)
      continue
    _, _, step = eval_lib.evaluate(bundle, dataset, use_graph=not args.no_graph, summaries=summaries)
    if config.eval_only_once:
      break
    num_steps = config.early_exit_steps if config.early_exit_steps is not None else config.max_steps
    if int(step) >= num_steps:
      break
    last_step = step


if __name__ == '__main__':
  main()
