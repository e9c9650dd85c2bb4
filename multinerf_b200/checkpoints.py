"""Checkpoint / resume of the TrainState (reference: flax.training.checkpoints as used by
train.py:84,219-223,284-287; eval.py:73; render.py:112).

State = {step, params, opt_state(count, mu, nu)} (SURVEY.md "Checkpoint / resume").  A checkpoint is
one `torch.save` file `checkpoint_<step>` holding the flat fp32 buffers, the layout they belong to
and the parameters as the reference's flax-named tree (`NerfMLP_0/Dense_3/{kernel,bias}`), so a
converted flax checkpoint can be loaded with `Model.init(flax_params=...)` and ours can be read
without this package.  Only rank 0 writes (train.py:216-223 `jax.host_id() == 0`).
"""
import os
import re

import torch

PREFIX = 'checkpoint_'


def _steps(ckpt_dir):
  if not ckpt_dir or not os.path.isdir(ckpt_dir):
    return []
  out = []
  for f in os.listdir(ckpt_dir):
    m = re.fullmatch(PREFIX + r'(\d+)', f)
    if m:
      out.append(int(m.group(1)))
  return sorted(out)


def latest_checkpoint(ckpt_dir):
  steps = _steps(ckpt_dir)
  return os.path.join(ckpt_dir, f'{PREFIX}{steps[-1]}') if steps else None


def save_checkpoint(ckpt_dir, state, step, keep=100, model=None):
  """Writes `checkpoint_<step>` atomically and keeps the `keep` newest files."""
  os.makedirs(ckpt_dir, exist_ok=True)
  p = state.params
  blob = {
      'step': int(step),
      'layout': {k: tuple(int(x) for x in v) for k, v in p.offsets.items()},
      'params': p.flat.detach().cpu(),
      'opt_state': {'count': int(p.step), 'mu': p.mu.detach().cpu(), 'nu': p.nu.detach().cpu()},
  }
  if model is not None:
    # tensors only (no numpy objects), so the file loads with weights_only=True
    blob['params_tree'] = {m: {l: {k: torch.from_numpy(v.copy()) for k, v in leaf.items()}
                               for l, leaf in layers.items()} if m in model.plans else
                           {k: torch.from_numpy(v.copy()) for k, v in layers.items()}
                           for m, layers in model.export_flax().items()}
  path = os.path.join(ckpt_dir, f'{PREFIX}{int(step)}')
  tmp = path + '.tmp'
  torch.save(blob, tmp)
  os.replace(tmp, path)
  for s in _steps(ckpt_dir)[:-keep] if keep else []:
    os.remove(os.path.join(ckpt_dir, f'{PREFIX}{s}'))
  return path


def restore_checkpoint(ckpt_dir, state, step=None, model=None):
  """Loads the newest (or the given) checkpoint into `state` in place; returns `state` unchanged
  when the directory holds none (flax semantics, train.py:84)."""
  path = latest_checkpoint(ckpt_dir) if step is None else os.path.join(ckpt_dir, f'{PREFIX}{int(step)}')
  if path is None or not os.path.exists(path):
    return state
  blob = torch.load(path, map_location='cpu', weights_only=True)   # tensors / ints / tuples only: no pickle code paths
  p = state.params
  layout = {k: tuple(v) for k, v in p.offsets.items()}
  if blob['layout'] != layout:
    raise ValueError(f'{path}: parameter layout differs from the model being restored '
                     f'({sorted(blob["layout"])} vs {sorted(layout)})')
  p.flat.copy_(blob['params'])
  p.mu.copy_(blob['opt_state']['mu'])
  p.nu.copy_(blob['opt_state']['nu'])
  p.step = int(blob['opt_state']['count'])
  if model is not None:
    model.bind(p)
    for mlp in model.mlps.values():
      mlp.repack()                      # refresh the bf16 operand copies of the weights
  return state
