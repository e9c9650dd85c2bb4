"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES.

Run in the build container only (needs /root/reference, which the GPU box lacks):
    python tests/golden/make_golden.py
JAX/Flax/gin are not installed, so the reference modules are imported with the
stand-ins under tests/golden/standin/ on sys.path: `jax.numpy` is numpy with 32-bit
result types, `jax.linearize`/`value_and_grad` are fp64 central differences, random
streams are explicit (see standin/README.md).  Inputs are seeded; every fixture stores
its inputs next to the reference's outputs so the tests need nothing but the .npz.

What this pins: the oracle's restatement of stepfun / render / coord / math /
ref_utils / geopoly / image (L2), and -- through the flax stand-in -- the real
`Model.__call__` / `MLP.__call__` control flow of internal/models.py.
What it cannot pin: XLA's own float rounding, threefry streams, flax initialisers,
optax.  Those stay "parity unpinned" (DESIGN.md).
"""
import math
import os
import sys
import types
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'standin'))
sys.path.insert(0, '/root/reference')
np.math = math  # numpy>=2 dropped the alias internal/ref_utils.py:55,72-81 relies on

for missing in ['dm_pix', 'cv2', 'rawpy', 'mediapy', 'optax', 'pycolmap', 'matplotlib',
                'tensorflow']:
  try:
    __import__(missing)
  except Exception:  # pylint: disable=broad-except
    sys.modules[missing] = mock.MagicMock()

import jax  # noqa: E402  (the stand-in)
import jax.numpy as jnp  # noqa: E402
from internal import coord, geopoly, image, ref_utils, render, stepfun  # noqa: E402
from internal import math as rmath  # noqa: E402

F = np.float32


def save(name, **arrays):
  path = os.path.join(HERE, name + '.npz')
  np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
  print(f'{name}.npz', {k: np.asarray(v).shape for k, v in arrays.items()})


def gen_geopoly():
  out = {}
  for shape in ['icosahedron', 'octahedron']:
    for v in [1, 2, 3, 4]:
      out[f'{shape}_{v}'] = geopoly.generate_basis(shape, v)
      out[f'{shape}_{v}_sym'] = geopoly.generate_basis(shape, v, remove_symmetries=False)
  save('geopoly', **out)


def gen_math():
  rng = np.random.default_rng(10)
  x = np.concatenate([rng.uniform(-400, 400, 2000), rng.uniform(-3e5, 3e5, 2000),
                      np.array([0., 100 * np.pi, -100 * np.pi, 314.15, 314.16, 1e10, -1e10])])
  x = x.astype(F)
  xe = np.concatenate([rng.uniform(-100, 100, 500), [87.9, 88.0, 88.1, 1e5, -1e5]]).astype(F)
  steps = np.array([0, 1, 10, 256, 511, 512, 513, 10000, 125000, 249999, 250000, 300000])
  lrs = np.array([float(rmath.learning_rate_decay(int(s), 2e-3, 2e-5, 250000, 512, 0.01))
                  for s in steps])
  lrs_nodelay = np.array([float(rmath.learning_rate_decay(int(s), 1e-3, 1e-5, 500000))
                          for s in steps])
  # interp family
  xp = np.sort(rng.uniform(0, 1, (16, 12)).astype(F), -1)
  xp[:, 0] = 0
  xp[:, -1] = 1
  xp[3, 4] = xp[3, 5]                       # a repeated knot
  fp = np.sort(rng.uniform(-2, 3, (16, 12)).astype(F), -1)
  xq = np.sort(rng.uniform(0, 1, (16, 7)).astype(F), -1)
  xq[0, 0] = 0.0
  save('math', x=x, safe_sin=rmath.safe_sin(x), safe_cos=rmath.safe_cos(x),
       xe=xe, safe_exp=rmath.safe_exp(xe), steps=steps, lrs=lrs, lrs_nodelay=lrs_nodelay,
       xp=xp, fp=fp, xq=xq, sorted_interp=rmath.sorted_interp(xq, xp, fp),
       interp=rmath.interp(xq, xp, fp))
  lin = rng.uniform(-0.1, 1.2, 1000).astype(F)
  lin[:5] = [0, 0.0031308, 0.0031309, 1.0, 0.5]
  save('image', linear=lin, srgb=image.linear_to_srgb(lin),
       mse=np.array([1e-4, 0.01, 0.5], F), psnr=image.mse_to_psnr(np.array([1e-4, 0.01, 0.5], F)))


def _rand_stepfun(rng, b, n, lo=0.0, hi=1.0, dup=False):
  t = np.sort(rng.uniform(lo, hi, (b, n + 1)).astype(F), -1)
  if dup:
    t[:, n // 2] = t[:, n // 2 - 1]         # an empty interval
  w = rng.uniform(0, 1, (b, n)).astype(F) ** 3
  w /= w.sum(-1, keepdims=True)
  return t, w.astype(F)


def gen_stepfun():
  rng = np.random.default_rng(20)
  out = {}
  # searchsorted / inner_outer / lossfun_outer
  a = np.sort(rng.uniform(0, 1, (8, 17)).astype(F), -1)
  v = rng.uniform(-0.2, 1.2, (8, 9)).astype(F)
  v[0, :3] = a[0, [0, 5, 16]]                # exact hits incl. both ends
  lo, hi = stepfun.searchsorted(a, v)
  out.update(ss_a=a, ss_v=v, ss_lo=lo, ss_hi=hi)
  t, w = _rand_stepfun(rng, 8, 12)
  te, we = _rand_stepfun(rng, 8, 20)
  te[:, 0] = 0
  te[:, -1] = 1
  inner, outer = stepfun.inner_outer(t, te, we)
  out.update(io_t=t, io_te=te, io_we=we, io_inner=inner, io_outer=outer,
             lo_w=w, lo_loss=stepfun.lossfun_outer(t, w, te, we))
  # max_dilate(_weights) at the two 360.gin dilation values and a big one
  t, w = _rand_stepfun(rng, 16, 64, dup=True)
  t[:, 0] = 0
  t[:, -1] = 1
  for tag, d in [('l1', 0.0103125), ('l2', 0.0026220703125), ('big', 0.3)]:
    td, wd = stepfun.max_dilate_weights(t, w, d, domain=(0.0, 1.0), renormalize=True)
    td2, pd2 = stepfun.max_dilate(t, stepfun.weight_to_pdf(t, w), d, domain=(0.0, 1.0))
    out.update({f'md_{tag}_t': td, f'md_{tag}_w': wd, f'md_{tag}_p': pd2})
  out.update(md_in_t=t, md_in_w=w)
  # integrate_weights / invert_cdf / sample / sample_intervals
  t, w = _rand_stepfun(rng, 16, 30, dup=True)
  logits = np.where(t[:, 1:] > t[:, :-1], 0.7 * np.log(w), -np.inf).astype(F)
  out.update(iw_w=w, iw_cw=stepfun.integrate_weights(w), si_t=t, si_logits=logits)
  u = np.sort(rng.uniform(0, 1, (16, 11)).astype(F), -1)
  out.update(ic_u=u, ic_t=stepfun.invert_cdf(u, t, logits),
             ic_t_gpu=stepfun.invert_cdf(u, t, logits, use_gpu_resampling=True))
  for ns in [8, 32]:
    out[f's_det_{ns}'] = stepfun.sample(None, t, logits, ns)
    out[f's_detc_{ns}'] = stepfun.sample(None, t, logits, ns, deterministic_center=True)
    out[f'si_det_{ns}'] = stepfun.sample_intervals(None, t, logits, ns, domain=(0.0, 1.0))
    for sj in [True, False]:
      jit = rng.uniform(0, 1, (16, 1 if sj else ns)).astype(F)
      key = jax.random.Stream([jit])
      out[f'si_jit_{ns}_{int(sj)}_in'] = jit
      out[f'si_jit_{ns}_{int(sj)}'] = stepfun.sample_intervals(
          key, t, logits, ns, single_jitter=sj, domain=(0.0, 1.0))
  # the reference's own known answer (tests/stepfun_test.py:579-586)
  out['si_single'] = stepfun.sample_intervals(
      None, np.array([3., 4.], F), np.array([0.], F), 10, single_jitter=False)
  # distortion + percentile
  t, w = _rand_stepfun(rng, 16, 32)
  w *= rng.uniform(0.2, 1.0, (16, 1)).astype(F)
  out.update(dl_t=t, dl_w=w, dl_loss=stepfun.lossfun_distortion(t, w))
  t, w = _rand_stepfun(rng, 16, 33, lo=0.2, hi=50.0)
  out.update(wp_t=t, wp_w=w, wp=stepfun.weighted_percentile(t, w, [5, 50, 95]))
  save('stepfun', **out)


def _rand_rays(rng, b):
  o = rng.uniform(-1, 1, (b, 3)).astype(F)
  d = rng.normal(size=(b, 3)).astype(F)
  d = d / np.linalg.norm(d, axis=-1, keepdims=True) * rng.uniform(0.8, 1.2, (b, 1)).astype(F)
  radii = rng.uniform(5e-4, 1e-3, (b, 1)).astype(F)
  return o, d.astype(F), radii


def gen_render():
  rng = np.random.default_rng(30)
  out = {}
  b, s = 12, 16
  o, d, radii = _rand_rays(rng, b)
  tdist = np.sort(rng.uniform(0.2, 30.0, (b, s + 1)).astype(F), -1)
  tdist[0, 3] = tdist[0, 2]                  # zero-width interval
  out.update(o=o, d=d, radii=radii, tdist=tdist)
  for shape in ['cone', 'cylinder']:
    for diag in [False, True]:
      m, c = render.cast_rays(tdist, o, d, radii, shape, diag=diag)
      out[f'cast_{shape}_{int(diag)}_mean'] = m
      out[f'cast_{shape}_{int(diag)}_cov'] = c
  density = (rng.uniform(0, 1, (b, s)) ** 4 * 20).astype(F)
  density[1] = 0
  density[2] = 1e4
  for ob in [False, True]:
    w, a, tr = render.compute_alpha_weights(density, tdist, d, opaque_background=ob)
    out.update({f'aw_{int(ob)}_w': w, f'aw_{int(ob)}_alpha': a, f'aw_{int(ob)}_trans': tr})
  rgbs = rng.uniform(0, 1, (b, s, 3)).astype(F)
  normals = rng.normal(size=(b, s, 3)).astype(F)
  rough = rng.uniform(0, 1, (b, s, 1)).astype(F)
  far = np.full((b, 1), 1e6, F)
  out.update(density=density, rgbs=rgbs, normals=normals, rough=rough, far=far)
  for ob in [False, True]:
    w = out[f'aw_{int(ob)}_w']
    r = render.volumetric_rendering(rgbs, w, tdist, 1.0, far, True,
                                    extras={'normals': normals, 'roughness': rough,
                                            'normals_pred': None})
    for k, v in r.items():
      out[f'vr_{int(ob)}_{k}'] = v
  bg = rng.uniform(0, 1, (b, 3)).astype(F)
  out['bg'] = bg
  out['vr_bg_rgb'] = render.volumetric_rendering(rgbs, out['aw_0_w'], tdist, bg, far, False)['rgb']
  save('render', **out)


def gen_coord():
  rng = np.random.default_rng(40)
  out = {}
  x = (rng.normal(size=(64, 3)) * np.array([0.3, 1.0, 8.0])[None]).astype(F)
  x[0] = 0
  x[1] = [1, 0, 0]
  x[2] = [0.6, 0.8, 0.0]
  a = rng.normal(size=(64, 3, 3)).astype(F) * 0.05
  cov = (a @ a.transpose(0, 2, 1) + 1e-4 * np.eye(3, dtype=F)).astype(F)
  zm, zc = coord.track_linearize(coord.contract, x, cov)
  out.update(x=x, cov=cov, contract=coord.contract(x), tl_mean=zm, tl_cov=zc)
  s = np.sort(rng.uniform(0, 1, (8, 9)).astype(F), -1)
  s[:, 0] = 0
  s[:, -1] = 1
  out['s'] = s
  for name, fn, near, far in [('none', None, 2.0, 6.0), ('reciprocal', jnp.reciprocal, 0.2, 1e6),
                              ('piecewise', 'piecewise', 0.0, 50.0), ('log', jnp.log, 0.5, 100.0)]:
    tn = np.full((8, 1), near, F)
    tf = np.full((8, 1), far, F)
    t_to_s, s_to_t = coord.construct_ray_warps(fn, tn, tf)
    tt = s_to_t(s)
    out[f'warp_{name}_t'] = tt
    out[f'warp_{name}_s'] = t_to_s(tt)
  for tag, shape, sub, mind, maxd in [('ico', 'icosahedron', 2, 0, 12), ('oct', 'octahedron', 1, 0, 16)]:
    basis = geopoly.generate_basis(shape, sub).astype(F)
    lm, lv = coord.lift_and_diagonalize(zm, zc, basis.T)
    enc = coord.integrated_pos_enc(lm, lv, mind, maxd)
    out.update({f'lift_{tag}_mean': lm, f'lift_{tag}_var': lv, f'ipe_{tag}': enc})
  dirs = rng.normal(size=(32, 3)).astype(F)
  dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
  out.update(dirs=dirs, pos_enc=coord.pos_enc(dirs, 0, 4, append_identity=True))
  nrm = rng.normal(size=(32, 3)).astype(F)
  nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
  out.update(nrm=nrm, reflect=ref_utils.reflect(dirs, nrm),
             l2n=ref_utils.l2_normalize(x[:32]))
  kinv = rng.uniform(0, 2, (32, 1)).astype(F)
  out.update(kinv=kinv, ide5=ref_utils.generate_ide_fn(5)(dirs, kinv),
             ide4=ref_utils.generate_ide_fn(4)(dirs, kinv),
             ide5_zero=ref_utils.generate_ide_fn(5)(dirs, np.zeros_like(kinv)))
  save('coord', **out)


if __name__ == '__main__':
  gen_geopoly()
  gen_math()
  gen_stepfun()
  gen_render()
  gen_coord()
  if os.path.exists(os.path.join(HERE, 'make_golden_model.py')):
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_golden_model',
                                                  os.path.join(HERE, 'make_golden_model.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main(save)
