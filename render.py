#!/usr/bin/env python
"""Offline rendering entry point (counterpart of the reference's render.py): test views or, with
`Config.render_path = True`, the dataset's camera path, from the newest checkpoint.

  python render.py --gin_configs=configs/360.gin --gin_bindings="Config.data_dir = '...'" \
      --gin_bindings="Config.checkpoint_dir = '...'" --gin_bindings="Config.render_path = True"
Frames go to <render_dir or checkpoint_dir/render>/{test_preds,path_renders}_step_<step>/.
"""
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from multinerf_b200 import configs, datasets, eval_lib  # noqa: E402
from train import parse, setup_distributed  # noqa: E402


def main(argv=None):
  args = parse(argv)
  world, rank, device = setup_distributed()
  bundle = configs.load_config(args.gin_configs, args.gin_bindings, search_paths=[ROOT, os.getcwd()])
  config = bundle.config
  dataset = datasets.load_dataset('test', config.data_dir, config, device=device, rank=rank, world=world)
  eval_lib.render(bundle, dataset, use_graph=not args.no_graph)


if __name__ == '__main__':
  main()
