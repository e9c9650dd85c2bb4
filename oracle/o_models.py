"""Oracle (test infrastructure): Model / MLP forward, torch-CPU, autograd-capable.

Follows /root/reference/internal/models.py:
  Model.__call__ :75-312 (level loop)     MLP.__call__ :402-612
Parameters use the flax tree of the reference:
  params[<'NerfMLP_0'|'PropMLP_0'>]['Dense_k'] = {'kernel': [in, out], 'bias': [out]}
in Dense creation order (models.py:455-460, 495, 515, 518, 521, 527, 577, 585), plus
params['exposure_scaling_offsets']['embedding'] / params['Embed_0']['embedding'].

Randomness is explicit: `rand` is None (rng=None in the reference) or a dict
  {'jitter': [per-level raw U[0,1) tensors], 'density_noise': [per-level N(0,1)],
   'bottleneck_noise': [per-level N(0,1)], 'bg': [per-level U[0,1) [B,3]]}.

`bf16=True` evaluates every Dense layer with bf16-rounded weights and bf16-rounded
layer inputs and fp32 accumulation -- the arithmetic of the tensor-core path.
PARITY UNPINNED at this level except through tests/golden (see oracle/__init__.py).
"""
import math

import torch
import torch.nn.functional as F

from . import o_coord
from . import o_math
from . import o_render
from . import o_stepfun


def _act(name):
  return {
      'relu': torch.relu, 'softplus': F.softplus, 'sigmoid': torch.sigmoid,
      'silu': F.silu, 'safe_exp': o_math.safe_exp, 'exp': torch.exp,
  }[name]


class _Dense:
  """Hands out Dense_k layers in creation order, like flax's auto-naming."""

  def __init__(self, tree, bf16):
    self.tree, self.k, self.bf16 = tree, 0, bf16

  def __call__(self, x):
    p = self.tree[f'Dense_{self.k}']
    self.k += 1
    w, b = p['kernel'], p['bias']
    if x.shape[-1] != w.shape[0]:
      raise ValueError(f'Dense_{self.k - 1}: input {x.shape[-1]} vs kernel {tuple(w.shape)}')
    if self.bf16:
      x = x.to(torch.bfloat16).to(torch.float32)
      w = w.to(torch.bfloat16).to(torch.float32)
    return x @ w + b


def mlp_apply(tree, cfg, basis, gaussians, viewdirs=None, glo_vec=None, rand=None,
              bf16=False, return_features=False):
  """MLP.__call__ (models.py:402-612).  basis: [K,3] (pos_basis_t is its transpose)."""
  dense = _Dense(tree, bf16)
  act = _act(cfg.net_activation)
  pos_basis_t = torch.as_tensor(basis, dtype=torch.float32).T.contiguous()
  means, covs = gaussians
  rand = rand or {}

  def predict_density(means, covs):
    if cfg.warp_fn is not None:
      assert cfg.warp_fn == 'contract'
      means, covs = o_coord.track_linearize_contract(means, covs)
    lifted_means, lifted_vars = o_coord.lift_and_diagonalize(means, covs, pos_basis_t)
    x = o_coord.integrated_pos_enc(lifted_means, lifted_vars, cfg.min_deg_point,
                                   cfg.max_deg_point)
    inputs = x
    for i in range(cfg.net_depth):
      x = act(dense(x))
      if i % cfg.skip_layer == 0 and i > 0:
        x = torch.cat([x, inputs], dim=-1)
    raw_density = dense(x)[..., 0]
    if rand.get('density_noise') is not None and cfg.density_noise > 0:
      raw_density = raw_density + cfg.density_noise * rand['density_noise']
    return raw_density, x, inputs

  if cfg.disable_density_normals:
    raw_density, x, feats = predict_density(means, covs)
    raw_grad_density, normals = None, None
  else:
    means_req = means.detach().requires_grad_(True) if not means.requires_grad else means
    raw_density, x, feats = predict_density(means_req, covs)
    (raw_grad_density,) = torch.autograd.grad(raw_density.sum(), means_req, create_graph=True)
    normals = -o_coord.l2_normalize(raw_grad_density)

  if cfg.enable_pred_normals:
    grad_pred = dense(x)
    normals_pred = -o_coord.l2_normalize(grad_pred)
    normals_to_use = normals_pred
  else:
    grad_pred, normals_pred, normals_to_use = None, None, normals

  density = _act(cfg.density_activation)(raw_density + cfg.density_bias)

  roughness = None
  if cfg.disable_rgb:
    rgb = torch.zeros_like(means)
  else:
    if viewdirs is not None:
      if cfg.use_diffuse_color:
        raw_rgb_diffuse = dense(x)
      if cfg.use_specular_tint:
        tint = torch.sigmoid(dense(x))
      if cfg.enable_pred_roughness:
        raw_roughness = dense(x)
        roughness = _act(cfg.roughness_activation)(raw_roughness + cfg.roughness_bias)
      if cfg.bottleneck_width > 0:
        bottleneck = dense(x)
        if rand.get('bottleneck_noise') is not None and cfg.bottleneck_noise > 0:
          bottleneck = bottleneck + cfg.bottleneck_noise * rand['bottleneck_noise']
        x = [bottleneck]
      else:
        x = []
      if cfg.use_directional_enc:
        dir_enc_fn = o_coord.generate_ide_fn(cfg.deg_view)
      else:
        dir_enc_fn = lambda d, _: o_coord.pos_enc(d, 0, cfg.deg_view, append_identity=True)
      if cfg.use_reflections:
        refdirs = o_coord.reflect(-viewdirs[..., None, :], normals_to_use)
        dir_enc = dir_enc_fn(refdirs, roughness)
      else:
        dir_enc = dir_enc_fn(viewdirs, roughness)
        dir_enc = dir_enc[..., None, :].expand(*bottleneck.shape[:-1], dir_enc.shape[-1])
      x.append(dir_enc)
      if cfg.use_n_dot_v:
        x.append((normals_to_use * viewdirs[..., None, :]).sum(dim=-1, keepdim=True))
      if glo_vec is not None:
        x.append(glo_vec[..., None, :].expand(*bottleneck.shape[:-1], glo_vec.shape[-1]))
      x = torch.cat(x, dim=-1)
      inputs = x
      for i in range(cfg.net_depth_viewdirs):
        x = act(dense(x))
        if i % cfg.skip_layer_dir == 0 and i > 0:
          x = torch.cat([x, inputs], dim=-1)
    rgb = _act(cfg.rgb_activation)(cfg.rgb_premultiplier * dense(x) + cfg.rgb_bias)
    if cfg.use_diffuse_color:
      diffuse_linear = torch.sigmoid(raw_rgb_diffuse - math.log(3.0))
      specular_linear = tint * rgb if cfg.use_specular_tint else 0.5 * rgb
      rgb = torch.clamp(o_math.linear_to_srgb(specular_linear + diffuse_linear), 0.0, 1.0)
    rgb = rgb * (1 + 2 * cfg.rgb_padding) - cfg.rgb_padding

  out = dict(density=density, rgb=rgb, raw_grad_density=raw_grad_density, grad_pred=grad_pred,
             normals=normals, normals_pred=normals_pred, roughness=roughness)
  if return_features:
    out['features'] = feats
    out['raw_density'] = raw_density
  return out


def level_schedule(mcfg, train_frac):
  """Per-level (num_samples, dilation, use_dilation, anneal) of models.py:147-179."""
  init_s_near = 0.0
  if mcfg.near_anneal_rate is not None:
    init_s_near = min(max(1 - train_frac / mcfg.near_anneal_rate, 0.0), mcfg.near_anneal_init)
  init_s_far = 1.0
  prod = 1
  out = []
  for i_level in range(mcfg.num_levels):
    is_prop = i_level < mcfg.num_levels - 1
    ns = mcfg.num_prop_samples if is_prop else mcfg.num_nerf_samples
    dilation = mcfg.dilation_bias + mcfg.dilation_multiplier * (init_s_far - init_s_near) / prod
    prod *= ns
    use_dilation = (mcfg.dilation_bias > 0 or mcfg.dilation_multiplier > 0) and i_level > 0
    if mcfg.anneal_slope > 0:
      s = mcfg.anneal_slope
      anneal = (s * train_frac) / ((s - 1) * train_frac + 1)
    else:
      anneal = 1.0
    out.append(dict(is_prop=is_prop, num_samples=ns, dilation=dilation,
                    use_dilation=use_dilation, anneal=anneal))
  return init_s_near, init_s_far, out


def model_apply(params, bundle, bases, rays, train_frac, compute_extras, rand=None,
                zero_glo=True, bf16=False, return_debug=False):
  """Model.__call__ (models.py:75-312).

  bases: {'nerf': [K,3], 'prop': [K,3]} projection bases.  rays: object with the
  attributes of utils.Rays (torch tensors, leading dims arbitrary).
  """
  mcfg, config = bundle.model, bundle.config
  nerf_cfg = bundle.nerf_mlp
  prop_cfg = nerf_cfg if mcfg.single_mlp else bundle.prop_mlp
  nerf_tree = params['NerfMLP_0']
  prop_tree = nerf_tree if mcfg.single_mlp else params.get('PropMLP_0')   # unused at num_levels == 1
  rand = rand or {}

  if mcfg.num_glo_features > 0:
    if not zero_glo:
      glo_vec = params['Embed_0']['embedding'][rays.cam_idx[..., 0].long()]
    else:
      glo_vec = torch.zeros(rays.origins.shape[:-1] + (mcfg.num_glo_features,))
  else:
    glo_vec = None

  _, s_to_t = o_coord.construct_ray_warps(mcfg.raydist_fn, rays.near, rays.far)
  init_s_near, init_s_far, sched = level_schedule(mcfg, train_frac)
  sdist = torch.cat([torch.full_like(rays.near, init_s_near),
                     torch.full_like(rays.far, init_s_far)], dim=-1)
  weights = torch.ones_like(rays.near)

  ray_history, renderings, debug = [], [], []
  for i_level, lv in enumerate(sched):
    if lv['use_dilation']:
      sdist, weights = o_stepfun.max_dilate_weights(
          sdist, weights, lv['dilation'], domain=(init_s_near, init_s_far), renormalize=True)
      sdist = sdist[..., 1:-1]
      weights = weights[..., 1:-1]
    logits = torch.where(sdist[..., 1:] > sdist[..., :-1],
                         lv['anneal'] * torch.log(weights + mcfg.resample_padding),
                         torch.tensor(-float('inf')))
    jit = rand['jitter'][i_level] if rand.get('jitter') is not None else None
    sdist, idx, cw = o_stepfun.sample_intervals(
        jit, sdist, logits, lv['num_samples'], single_jitter=mcfg.single_jitter,
        domain=(init_s_near, init_s_far), use_gpu_resampling=mcfg.use_gpu_resampling,
        return_index=True)
    if mcfg.stop_level_grad:
      sdist = sdist.detach()
    tdist = s_to_t(sdist)
    gaussians = o_render.cast_rays(tdist, rays.origins, rays.directions, rays.radii,
                                   mcfg.ray_shape, diag=False)
    if mcfg.disable_integration:
      gaussians = (gaussians[0], torch.zeros_like(gaussians[1]))
    is_prop = lv['is_prop']
    mlp_rand = {
        'density_noise': rand['density_noise'][i_level] if rand.get('density_noise') else None,
        'bottleneck_noise': (rand['bottleneck_noise'][i_level]
                             if rand.get('bottleneck_noise') else None),
    }
    res = mlp_apply(prop_tree if is_prop else nerf_tree, prop_cfg if is_prop else nerf_cfg,
                    bases['prop'] if is_prop else bases['nerf'], gaussians,
                    viewdirs=rays.viewdirs if mcfg.use_viewdirs else None,
                    glo_vec=None if is_prop else glo_vec, rand=mlp_rand, bf16=bf16,
                    return_features=return_debug)
    weights = o_render.compute_alpha_weights(res['density'], tdist, rays.directions,
                                             opaque_background=mcfg.opaque_background)[0]
    lo, hi = mcfg.bg_intensity_range
    if lo == hi:
      bg_rgbs = lo
    elif rand.get('bg') is None:
      bg_rgbs = (lo + hi) / 2
    else:
      bg_rgbs = lo + (hi - lo) * rand['bg'][i_level]

    if getattr(rays, 'exposure_idx', None) is not None:
      res['rgb'] = res['rgb'] * rays.exposure_values[..., None, :]
      if mcfg.learned_exposure_scaling:
        eidx = rays.exposure_idx[..., 0].long()
        mask = (eidx > 0).to(torch.float32)
        scaling = 1 + mask[..., None] * params['exposure_scaling_offsets']['embedding'][eidx]
        res['rgb'] = res['rgb'] * scaling[..., None, :]

    rendering = o_render.volumetric_rendering(
        res['rgb'], weights, tdist, bg_rgbs, rays.far, compute_extras,
        extras={k: v for k, v in res.items() if k.startswith('normals') or k in ['roughness']})
    if compute_extras:
      n = config.vis_num_rays
      rendering['ray_sdist'] = sdist.reshape(-1, sdist.shape[-1])[:n, :]
      rendering['ray_weights'] = weights.reshape(-1, weights.shape[-1])[:n, :]
      rgb = res['rgb']
      rendering['ray_rgbs'] = rgb.reshape((-1,) + rgb.shape[-2:])[:n, :, :]
    renderings.append(rendering)
    res['sdist'] = sdist.clone()
    res['weights'] = weights.clone()
    ray_history.append(res)
    debug.append(dict(idx=idx, cw=cw, tdist=tdist))

  if compute_extras:
    ws = [r['ray_weights'] for r in renderings]
    rgbs = [r['ray_rgbs'] for r in renderings]
    final_rgb = (rgbs[-1] * ws[-1][..., None]).sum(dim=-2)
    for i in range(len(rgbs) - 1):
      renderings[i]['ray_rgbs'] = final_rgb[:, None, :].expand(rgbs[i].shape)
  if return_debug:
    return renderings, ray_history, debug
  return renderings, ray_history
