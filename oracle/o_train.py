"""Oracle (test infrastructure): losses, gradient clipping, Adam, one train step.

Follows /root/reference/internal/train_utils.py:
  compute_data_loss :72-136   interlevel_loss :139-150   distortion_loss :153-159
  orientation_loss :162-178   predicted_normal_loss :181-197
  clip_gradients :200-218     train_step :239-339        create_optimizer :349-374
optax.adam is not in /root/reference (requirements.txt:1-11, unpinned); it is restated
from its published definition: m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
m_hat = m/(1-b1^t); v_hat = v/(1-b2^t); update = -lr(t-1) * m_hat/(sqrt(v_hat)+eps),
t = 1-based update count, schedule evaluated at the 0-based count.  PARITY UNPINNED.
"""
import math

import torch

from . import o_math
from . import o_models
from . import o_stepfun

EPS = o_math.EPS


def compute_data_loss(batch_rgb, renderings, lossmult, config):
  data_losses, mses = [], []
  lossmult = lossmult.expand_as(batch_rgb[..., :3])
  if config.disable_multiscale_loss:
    lossmult = torch.ones_like(lossmult)
  for rendering in renderings:
    resid_sq = (rendering['rgb'] - batch_rgb[..., :3]) ** 2
    denom = lossmult.sum()
    mses.append((lossmult * resid_sq).sum() / denom)
    if config.data_loss_type == 'mse':
      data_loss = resid_sq
    elif config.data_loss_type == 'charb':
      data_loss = torch.sqrt(resid_sq + config.charb_padding ** 2)
    elif config.data_loss_type == 'rawnerf':
      clip = torch.clamp(rendering['rgb'], max=1.0)
      resid_sq_clip = (clip - batch_rgb[..., :3]) ** 2
      scaling_grad = 1.0 / (1e-3 + clip.detach())
      data_loss = resid_sq_clip * scaling_grad ** 2
    else:
      raise AssertionError(config.data_loss_type)
    data_losses.append((lossmult * data_loss).sum() / denom)
  data_losses = torch.stack(data_losses)
  loss = (config.data_coarse_loss_mult * data_losses[:-1].sum() +
          config.data_loss_mult * data_losses[-1])
  return loss, {'mses': torch.stack(mses)}


def interlevel_loss(ray_history, config):
  c = ray_history[-1]['sdist'].detach()
  w = ray_history[-1]['weights'].detach()
  total = 0.0
  for rr in ray_history[:-1]:
    total = total + o_stepfun.lossfun_outer(c, w, rr['sdist'], rr['weights']).mean()
  return config.interlevel_loss_mult * total


def distortion_loss(ray_history, config):
  c, w = ray_history[-1]['sdist'], ray_history[-1]['weights']
  return config.distortion_loss_mult * o_stepfun.lossfun_distortion(c, w).mean()


def orientation_loss(viewdirs, num_levels, ray_history, config):
  total = 0.0
  for i, rr in enumerate(ray_history):
    w = rr['weights']
    n = rr[config.orientation_loss_target]
    if n is None:
      raise ValueError('Normals cannot be None if orientation loss is on.')
    v = -1.0 * viewdirs
    n_dot_v = (n * v[..., None, :]).sum(dim=-1)
    loss = (w * torch.clamp(n_dot_v, max=0.0) ** 2).sum(dim=-1).mean()
    mult = config.orientation_coarse_loss_mult if i < num_levels - 1 else \
        config.orientation_loss_mult
    total = total + mult * loss
  return total


def predicted_normal_loss(num_levels, ray_history, config):
  total = 0.0
  for i, rr in enumerate(ray_history):
    w, n, n_pred = rr['weights'], rr['normals'], rr['normals_pred']
    if n is None or n_pred is None:
      raise ValueError('Predicted normals and gradient normals cannot be None if '
                       'predicted normal loss is on.')
    loss = (w * (1.0 - (n * n_pred).sum(dim=-1))).sum(dim=-1).mean()
    mult = config.predicted_normal_coarse_loss_mult if i < num_levels - 1 else \
        config.predicted_normal_loss_mult
    total = total + mult * loss
  return total


def loss_fn(params, bundle, bases, rays, batch_rgb, train_frac, rand=None, bf16=False):
  """The closure of train_utils.py:265-314; returns (loss, stats, aux)."""
  config = bundle.config
  renderings, ray_history = o_models.model_apply(
      params, bundle, bases, rays, train_frac, compute_extras=False, rand=rand,
      zero_glo=False, bf16=bf16)
  losses = {}
  losses['data'], stats = compute_data_loss(batch_rgb, renderings, rays.lossmult, config)
  if config.interlevel_loss_mult > 0:
    losses['interlevel'] = interlevel_loss(ray_history, config)
  if config.distortion_loss_mult > 0:
    losses['distortion'] = distortion_loss(ray_history, config)
  if config.orientation_coarse_loss_mult > 0 or config.orientation_loss_mult > 0:
    losses['orientation'] = orientation_loss(rays.viewdirs, bundle.model.num_levels,
                                             ray_history, config)
  if config.predicted_normal_coarse_loss_mult > 0 or config.predicted_normal_loss_mult > 0:
    losses['predicted_normals'] = predicted_normal_loss(bundle.model.num_levels, ray_history,
                                                        config)
  if getattr(config, 'weight_decay_mults', None):
    # train_utils.py:304-309: keys 'Module', 'Module/Dense_k' (or 'Module/Dense_k/kernel')
    def norm_sq(tree):
      if isinstance(tree, dict):
        return sum(norm_sq(v) for v in tree.values())
      return (tree ** 2).sum()
    total = 0.0
    for key, mult in dict(config.weight_decay_mults).items():
      sub = params
      for part in key.split('/'):
        sub = sub[part]
      total = total + mult * norm_sq(sub)
    losses['weight'] = total
  stats['losses'] = losses
  stats['loss'] = sum(losses.values())
  return stats['loss'], stats, (renderings, ray_history)


def _leaves(tree, prefix=()):
  for k in sorted(tree.keys()):
    v = tree[k]
    if isinstance(v, dict):
      yield from _leaves(v, prefix + (k,))
    else:
      yield prefix + (k,), v


def clip_gradients(grads, config):
  """Per top-level module: clip by value, then by global norm (train_utils.py:200-218)."""
  out = {}
  for k, sub in grads.items():
    leaves = dict(_leaves(sub)) if isinstance(sub, dict) else {(): sub}
    if config.grad_max_val > 0:
      leaves = {p: g.clamp(-config.grad_max_val, config.grad_max_val) for p, g in leaves.items()}
    if config.grad_max_norm > 0:
      norm = torch.sqrt(sum((g ** 2).sum() for g in leaves.values()))
      mult = torch.clamp(config.grad_max_norm / (EPS + norm), max=1.0)
      leaves = {p: mult * g for p, g in leaves.items()}
    out[k] = leaves
  return out


def lr_at(step, config):
  return o_math.learning_rate_decay(step, config.lr_init, config.lr_final, config.max_steps,
                                    config.lr_delay_steps, config.lr_delay_mult)


def adam_update(p, g, m, v, count, lr, config):
  """One optax.adam update; `count` is the 0-based number of updates already applied."""
  b1, b2, eps = config.adam_beta1, config.adam_beta2, config.adam_eps
  m = b1 * m + (1 - b1) * g
  v = b2 * v + (1 - b2) * g * g
  t = count + 1
  m_hat = m / (1 - b1 ** t)
  v_hat = v / (1 - b2 ** t)
  return p - lr * m_hat / (torch.sqrt(v_hat) + eps), m, v


def train_step(params, opt_state, bundle, bases, rays, batch_rgb, train_frac, rand=None,
               bf16=False):
  """train_utils.py:239-339 for one device.  params: nested dict of tensors.
  opt_state: {'count': int, 'mu': tree, 'nu': tree} (zeros at start).
  Returns (new_params, new_opt_state, stats, raw_grads)."""
  config = bundle.config
  flat = {}
  for top, sub in params.items():
    for path, leaf in _leaves(sub):
      flat[(top,) + path] = leaf.detach().clone().requires_grad_(True)

  def unflatten(fl):
    tree = {}
    for path, leaf in fl.items():
      d = tree
      for k in path[:-1]:
        d = d.setdefault(k, {})
      d[path[-1]] = leaf
    return tree

  loss, stats, _ = loss_fn(unflatten(flat), bundle, bases, rays, batch_rgb, train_frac, rand,
                           bf16)
  keys = list(flat.keys())
  grads = torch.autograd.grad(loss, [flat[k] for k in keys], allow_unused=True)
  gflat = {k: (g if g is not None else torch.zeros_like(flat[k])) for k, g in zip(keys, grads)}
  gtree = {}
  for k, g in gflat.items():
    gtree.setdefault(k[0], {})[k[1:]] = g
  # clip per top-level module
  clipped = {}
  for top, leaves in gtree.items():
    if config.grad_max_val > 0:
      leaves = {p: g.clamp(-config.grad_max_val, config.grad_max_val) for p, g in leaves.items()}
    if config.grad_max_norm > 0:
      norm = torch.sqrt(sum((g ** 2).sum() for g in leaves.values()))
      mult = torch.clamp(config.grad_max_norm / (EPS + norm), max=1.0)
      leaves = {p: mult * g for p, g in leaves.items()}
    for p, g in leaves.items():
      clipped[(top,) + p] = torch.nan_to_num(g)      # train_utils.py:328
  count = opt_state['count']
  lr = lr_at(count, config)
  new_flat, mu, nu = {}, {}, {}
  for k in keys:
    m0 = opt_state['mu'].get(k, torch.zeros_like(flat[k]))
    v0 = opt_state['nu'].get(k, torch.zeros_like(flat[k]))
    new_flat[k], mu[k], nu[k] = adam_update(flat[k].detach(), clipped[k], m0, v0, count, lr,
                                            config)
  stats['psnrs'] = o_math.mse_to_psnr(stats['mses'])
  stats['psnr'] = stats['psnrs'][-1]
  new_state = {'count': count + 1, 'mu': mu, 'nu': nu}
  return unflatten(new_flat), new_state, stats, gflat
