"""Model-level golden vectors: runs the reference's real `Model.__call__` / `MLP.__call__`
(internal/models.py) and loss functions (internal/train_utils.py) under the jax/flax/gin stand-ins
(tests/golden/standin/).  Called from make_golden.py (`main(save)`), build container only.

Each fixture stores: the gin-style bindings, the rays, the parameter tree (flattened
'Module/Dense_k/kernel' keys), the explicit random draws, and the reference's outputs per level.
`jax.linearize` / `jax.value_and_grad` are fp64 central differences in the stand-in, so the
contraction Jacobian and the Ref-NeRF density normals carry ~1e-6 relative error.
"""
import numpy as np

import gin
import jax
from flax import linen as nn
from internal import configs as rconfigs
from internal import coord, models, train_utils, utils
from internal import math as rmath
import jax.numpy as jnp

F = np.float32


def _init_params(model, rng, rays):
  """Shape discovery: run once with auto-creating parameter dicts (he_uniform-like values)."""
  def make(shape_in, features):
    lim = np.sqrt(6.0 / shape_in)
    return {'kernel': rng.uniform(-lim, lim, (shape_in, features)).astype(F),
            'bias': (rng.normal(size=(features,)) * 0.1).astype(F)}
  orig_dense, orig_embed = nn.Dense.__call__, nn.Embed.__call__

  def dense_call(self, x):
    if 'kernel' not in self._params:
      self._params.update(make(x.shape[-1], self.features))
    return orig_dense(self, x)

  def embed_call(self, idx):
    if 'embedding' not in self._params:
      self._params['embedding'] = (rng.normal(size=(self.num_embeddings, self.features)) * 0.1).astype(F)
    return orig_embed(self, idx)
  nn.Dense.__call__, nn.Embed.__call__ = dense_call, embed_call
  try:
    tree = AutoDict()
    model.apply({'params': tree}, None, rays, train_frac=1.0, compute_extras=False, zero_glo=False)
  finally:
    nn.Dense.__call__, nn.Embed.__call__ = orig_dense, orig_embed
  return tree.plain()


class AutoDict(dict):
  def __missing__(self, k):
    v = AutoDict()
    self[k] = v
    return v

  def plain(self):
    return {k: (v.plain() if isinstance(v, AutoDict) else v) for k, v in self.items()}


def _flatten(tree, prefix=''):
  out = {}
  for k, v in tree.items():
    if isinstance(v, dict):
      out.update(_flatten(v, prefix + k + '/'))
    else:
      out[prefix + k] = np.asarray(v)
  return out


def _rays(rng, B, near, far, kind, exposure=False):
  if kind == 'cube':
    o = rng.uniform(-1, 1, (B, 3))
    d = rng.normal(size=(B, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
  elif kind == 'sphere':
    o = rng.normal(size=(B, 3))
    o = o / np.linalg.norm(o, axis=-1, keepdims=True) * 4.0
    d = -o / 4.0 + rng.normal(size=(B, 3)) * 0.1
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
  else:  # ndc-like
    o = np.concatenate([rng.uniform(-1, 1, (B, 2)), -np.ones((B, 1))], -1)
    d = np.concatenate([rng.uniform(-.5, .5, (B, 2)), 2 * np.ones((B, 1))], -1)
  v = d / np.linalg.norm(d, axis=-1, keepdims=True)
  d = d * rng.uniform(0.8, 1.2, (B, 1))
  kw = {}
  if exposure:
    eidx = rng.integers(0, 4, (B, 1)).astype(np.int32)
    kw = dict(exposure_idx=eidx, exposure_values=(2.0 ** -eidx).astype(F))
  lossmult = np.eye(3, dtype=F)[rng.integers(0, 3, B)] if exposure else np.ones((B, 1), F)
  return utils.Rays(origins=o.astype(F), directions=d.astype(F), viewdirs=v.astype(F),
                    radii=rng.uniform(5e-4, 1e-3, (B, 1)).astype(F), imageplane=np.zeros((B, 2), F),
                    lossmult=lossmult, near=np.full((B, 1), near, F), far=np.full((B, 1), far, F),
                    cam_idx=np.zeros((B, 1), np.int32), **kw)


CASES = {
    'mini360': dict(
        near=0.2, far=1e6, rays='cube', B=24, train_frac=0.5,
        Config=dict(),
        Model=dict(raydist_fn=jnp.reciprocal, opaque_background=True, num_prop_samples=16, num_nerf_samples=8),
        PropMLP=dict(warp_fn=coord.contract, net_depth=2, net_width=32, disable_density_normals=True,
                     disable_rgb=True),
        NerfMLP=dict(warp_fn=coord.contract, net_depth=6, net_width=48, bottleneck_width=16,
                     net_width_viewdirs=24, disable_density_normals=True)),
    'plumbing': dict(
        near=2.0, far=6.0, rays='sphere', B=16, train_frac=1.0,
        Config=dict(data_loss_type='mse', distortion_loss_mult=0.0),
        Model=dict(num_levels=1, num_nerf_samples=12),
        PropMLP=dict(net_depth=2, net_width=16, basis_shape='octahedron', basis_subdivisions=1,
                     disable_density_normals=True, disable_rgb=True, max_deg_point=16),
        NerfMLP=dict(net_depth=5, net_width=32, bottleneck_width=16, net_width_viewdirs=16,
                     basis_shape='octahedron', basis_subdivisions=1, disable_density_normals=True,
                     max_deg_point=16)),
    'miniraw': dict(
        near=0.0, far=1.0, rays='ndc', B=20, train_frac=0.3, exposure=True,
        Config=dict(data_loss_type='rawnerf', interlevel_loss_mult=0.0, distortion_loss_mult=0.01,
                    data_coarse_loss_mult=0.1, grad_max_norm=0.1, grad_max_val=0.1),
        Model=dict(ray_shape='cylinder', learned_exposure_scaling=True, num_levels=2, num_prop_samples=12,
                   num_nerf_samples=12, opaque_background=True, single_mlp=True, anneal_slope=0.,
                   dilation_multiplier=0., dilation_bias=0., single_jitter=False),
        PropMLP=dict(),
        NerfMLP=dict(net_depth=5, net_width=32, bottleneck_width=16, net_width_viewdirs=16,
                     basis_shape='octahedron', basis_subdivisions=1, disable_density_normals=True,
                     max_deg_point=16, rgb_padding=0., rgb_activation=rmath.safe_exp, rgb_bias=-5.,
                     density_noise=1.)),
    'miniglo': dict(        # GLO vectors, random backgrounds, bottleneck noise, near-plane annealing
        near=0.2, far=1e6, rays='cube', B=20, train_frac=0.2, cam_idx=6,
        Config=dict(),
        Model=dict(raydist_fn=jnp.reciprocal, num_prop_samples=12, num_nerf_samples=8, num_glo_features=4,
                   num_glo_embeddings=6, bg_intensity_range=(0.2, 0.9), near_anneal_rate=0.5,
                   near_anneal_init=0.9),
        PropMLP=dict(warp_fn=coord.contract, net_depth=2, net_width=32, disable_density_normals=True,
                     disable_rgb=True),
        NerfMLP=dict(warp_fn=coord.contract, net_depth=5, net_width=48, bottleneck_width=16,
                     net_width_viewdirs=24, disable_density_normals=True, bottleneck_noise=0.3)),
    'minirefnerf': dict(
        near=2.0, far=6.0, rays='sphere', B=12, train_frac=0.7,
        Config=dict(data_loss_type='mse', distortion_loss_mult=0.0, orientation_loss_mult=0.1,
                    orientation_loss_target='normals_pred', predicted_normal_loss_mult=3e-4,
                    orientation_coarse_loss_mult=0.01, predicted_normal_coarse_loss_mult=3e-5,
                    interlevel_loss_mult=0.0, data_coarse_loss_mult=0.1),
        Model=dict(num_levels=2, single_mlp=True, num_prop_samples=8, num_nerf_samples=8, anneal_slope=0.,
                   dilation_multiplier=0., dilation_bias=0., single_jitter=False, resample_padding=0.01),
        PropMLP=dict(),
        NerfMLP=dict(net_depth=5, net_width=32, net_depth_viewdirs=6, net_width_viewdirs=16,
                     basis_shape='octahedron', basis_subdivisions=1, disable_density_normals=False,
                     enable_pred_normals=True, use_directional_enc=True, use_reflections=True, deg_view=5,
                     enable_pred_roughness=True, use_diffuse_color=True, use_specular_tint=True,
                     use_n_dot_v=True, bottleneck_width=16, density_bias=0.5, max_deg_point=16)),
}


def _name(v):
  if callable(v):
    return getattr(v, '__name__', str(v))
  return v


def run_case(tag, spec, save):
  rng = np.random.default_rng(abs(hash(tag)) % 2 ** 31 if False else {'mini360': 1, 'plumbing': 2, 'miniraw': 3,
                                                                      'minirefnerf': 4, 'miniglo': 5}[tag])
  gin.clear()
  for cls in ['Model', 'PropMLP', 'NerfMLP']:
    gin.bind(cls, **spec[cls])
  config = rconfigs.Config(**spec['Config'])
  model = models.Model(config=config)
  B = spec['B']
  rays = _rays(rng, B, spec['near'], spec['far'], spec['rays'], spec.get('exposure', False))
  if spec.get('cam_idx'):
    import dataclasses
    rays = dataclasses.replace(rays, cam_idx=rng.integers(0, spec['cam_idx'], (B, 1)).astype(np.int32))
  params = _init_params(model, rng, rays)
  out = {'meta_tag': np.array(tag)}
  for cls in ['Config', 'Model', 'PropMLP', 'NerfMLP']:
    for k, v in spec[cls].items():
      out[f'bind/{cls}/{k}'] = np.array(_name(v))
  out.update({'meta_near': spec['near'], 'meta_far': spec['far'], 'meta_train_frac': spec['train_frac']})
  for f, v in rays.__dict__.items():
    if v is not None:
      out[f'rays/{f}'] = v
  out.update({'params/' + k: v for k, v in _flatten(params).items()})
  target = rng.uniform(0, 1, (B, 3)).astype(F)
  out['target'] = target
  num_levels = model.num_levels
  for mode in ['det', 'rand']:
    draws, names = [], []
    key = None
    if mode == 'rand':
      for lv in range(num_levels):
        is_prop = lv < num_levels - 1
        S = model.num_prop_samples if is_prop else model.num_nerf_samples
        j = rng.uniform(0, 1, (B, 1 if model.single_jitter else S)).astype(F)
        draws.append(j)
        out[f'{mode}/jitter{lv}'] = j
        mlp_bind = spec['NerfMLP'] if (model.single_mlp or not is_prop) else spec['PropMLP']
        if mlp_bind.get('density_noise', 0) > 0:
          nz = rng.normal(size=(B, S)).astype(F)
          draws.append(nz)
          out[f'{mode}/density_noise{lv}'] = nz
        if mlp_bind.get('bottleneck_noise', 0) > 0 and not mlp_bind.get('disable_rgb', False):
          bn = rng.normal(size=(B, S, mlp_bind.get('bottleneck_width', 256))).astype(F)   # models.py:529-533
          draws.append(bn)
          out[f'{mode}/bottleneck_noise{lv}'] = bn
        lo, hi = spec['Model'].get('bg_intensity_range', (1., 1.))
        if lo != hi:
          bg = rng.uniform(0, 1, (B, 3)).astype(F)                                         # models.py:249-254
          draws.append(bg)
          out[f'{mode}/bg{lv}'] = bg
      key = jax.random.Stream(draws)
    renderings, ray_history = model.apply({'params': params}, key, rays, train_frac=spec['train_frac'],
                                          compute_extras=True, zero_glo=False)
    if mode == 'rand':
      assert not key.draws, 'unconsumed random draws'
    for lv, (r, h) in enumerate(zip(renderings, ray_history)):
      for k, v in r.items():
        out[f'{mode}/rend{lv}/{k}'] = np.asarray(v)
      for k, v in h.items():
        if v is not None:
          out[f'{mode}/hist{lv}/{k}'] = np.asarray(v)
    # losses of the train-step closure on these outputs (train_utils.py:72-197)
    batch = utils.Batch(rays=rays, rgb=target)
    data_loss, stats = train_utils.compute_data_loss(batch, renderings, rays, 1.0, config)
    out[f'{mode}/loss_data'] = np.asarray(data_loss)
    out[f'{mode}/mses'] = np.asarray(stats['mses'])
    out[f'{mode}/loss_interlevel'] = np.asarray(train_utils.interlevel_loss(ray_history, config))
    out[f'{mode}/loss_distortion'] = np.asarray(train_utils.distortion_loss(ray_history, config))
    if config.orientation_loss_mult > 0:
      out[f'{mode}/loss_orientation'] = np.asarray(train_utils.orientation_loss(rays, model, ray_history, config))
      out[f'{mode}/loss_pred_normals'] = np.asarray(train_utils.predicted_normal_loss(model, ray_history, config))
  # clip_gradients on a synthetic gradient tree (train_utils.py:200-218)
  class FakeGrad(dict):
    pass
  g = {'params': {k: {kk: {n: (rng.normal(size=np.asarray(a).shape) * 0.05).astype(F) for n, a in vv.items()}
                      if isinstance(vv, dict) else None for kk, vv in v.items()} if 'embedding' not in v
                  else {'embedding': (rng.normal(size=np.asarray(v['embedding']).shape) * 0.05).astype(F)}
                  for k, v in params.items()}}
  clipped = train_utils.clip_gradients(FakeGrad(g), config)
  out.update({'clip_in/' + k: v for k, v in _flatten(g['params']).items()})
  out.update({'clip_out/' + k: np.asarray(v) for k, v in _flatten(clipped['params']).items()})
  save('model_' + tag, **out)


def main(save):
  for tag, spec in CASES.items():
    run_case(tag, spec, save)
