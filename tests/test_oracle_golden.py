"""Oracle vs golden vectors produced by running the reference's own source files
(tests/golden/make_golden.py), plus the known answers of the reference's unit tests.
CPU only."""
import math

import numpy as np
import pytest
import torch

from oracle import o_coord, o_math, o_render, o_stepfun
from multinerf_b200 import geopoly
from util import T, close, golden


def test_geopoly_matches_reference_run():
  g = golden('geopoly')
  for shape in ['icosahedron', 'octahedron']:
    for v in [1, 2, 3, 4]:
      np.testing.assert_array_equal(geopoly.generate_basis(shape, v), g[f'{shape}_{v}'])
      np.testing.assert_array_equal(
          geopoly.generate_basis(shape, v, remove_symmetries=False), g[f'{shape}_{v}_sym'])


def test_geopoly_reference_test_goldens():
  # tests/geopoly_test.py:78-101 (icosahedron/2 -> 21 vectors, rounded goldens, up to
  # sign/permutation) -- spot values from that table.
  b = geopoly.generate_basis('icosahedron', 2)
  assert b.shape == (21, 3)
  gold = np.array([[0.85065081, 0.00000000, 0.52573111], [0.80901699, 0.50000000, 0.30901699],
                   [0.52573111, 0.85065081, 0.00000000], [1.00000000, 0.00000000, 0.00000000]])
  for row in gold:
    d = np.minimum(np.abs(b - row).sum(-1), np.abs(b + row).sum(-1))
    assert d.min() < 1e-6
  assert generate_counts() == (6, 21, 3, 33)


def generate_counts():
  return (geopoly.generate_basis('icosahedron', 1).shape[0],
          geopoly.generate_basis('icosahedron', 2).shape[0],
          geopoly.generate_basis('octahedron', 1).shape[0],
          geopoly.generate_basis('octahedron', 4).shape[0])


def test_octahedron_1_is_minus_antidiagonal():
  np.testing.assert_array_equal(geopoly.generate_basis('octahedron', 1),
                                np.array([[0, 0, -1], [0, -1, 0], [-1, 0, 0]], np.float64))


def test_math_golden():
  g = golden('math')
  x = T(g['x'])
  close(o_math.safe_sin(x), g['safe_sin'], atol=2e-6, rtol=0, msg='safe_sin')
  close(o_math.safe_cos(x), g['safe_cos'], atol=2e-6, rtol=0, msg='safe_cos')
  close(o_math.safe_exp(T(g['xe'])), g['safe_exp'], atol=0, rtol=2e-6, msg='safe_exp')
  for s, lr, lr2 in zip(g['steps'], g['lrs'], g['lrs_nodelay']):
    assert abs(o_math.learning_rate_decay(int(s), 2e-3, 2e-5, 250000, 512, 0.01) - lr) <= 1e-6 * lr
    assert abs(o_math.learning_rate_decay(int(s), 1e-3, 1e-5, 500000) - lr2) <= 1e-6 * lr2
  close(o_math.sorted_interp(T(g['xq']), T(g['xp']), T(g['fp'])), g['sorted_interp'], msg='sorted_interp')
  close(o_math.interp(T(g['xq']), T(g['xp']), T(g['fp'])), g['interp'], msg='interp')
  gi = golden('image')
  close(o_math.linear_to_srgb(T(gi['linear'])), gi['srgb'], msg='srgb')
  close(o_math.mse_to_psnr(T(gi['mse'])), gi['psnr'], msg='psnr')


def test_safe_exp_grad_is_value():
  x = torch.tensor([-3.0, 0.0, 50.0, 100.0], requires_grad=True)
  y = o_math.safe_exp(x)
  y.sum().backward()
  close(x.grad, y.detach(), rtol=1e-6)
  assert torch.isfinite(x.grad).all()


def test_stepfun_golden():
  g = golden('stepfun')
  lo, hi = o_stepfun.searchsorted(T(g['ss_a']), T(g['ss_v']))
  np.testing.assert_array_equal(lo.numpy(), g['ss_lo'])
  np.testing.assert_array_equal(hi.numpy(), g['ss_hi'])
  inner, outer = o_stepfun.inner_outer(T(g['io_t']), T(g['io_te']), T(g['io_we']))
  close(inner, g['io_inner'], msg='inner')
  close(outer, g['io_outer'], msg='outer')
  close(o_stepfun.lossfun_outer(T(g['io_t']), T(g['lo_w']), T(g['io_te']), T(g['io_we'])),
        g['lo_loss'], msg='lossfun_outer')
  t, w = T(g['md_in_t']), T(g['md_in_w'])
  for tag, d in [('l1', 0.0103125), ('l2', 0.0026220703125), ('big', 0.3)]:
    td, wd = o_stepfun.max_dilate_weights(t, w, d, domain=(0.0, 1.0), renormalize=True)
    np.testing.assert_array_equal(td.numpy(), g[f'md_{tag}_t'])   # sort order: exact
    close(wd, g[f'md_{tag}_w'], atol=1e-7, rtol=1e-5, msg=f'max_dilate_weights {tag}')
    _, pd = o_stepfun.max_dilate(t, o_stepfun.weight_to_pdf(t, w), d, domain=(0.0, 1.0))
    close(pd, g[f'md_{tag}_p'], atol=0, rtol=1e-6, msg=f'max_dilate {tag}')
  close(o_stepfun.integrate_weights(T(g['iw_w'])), g['iw_cw'], atol=1e-6, msg='integrate_weights')
  t, logits = T(g['si_t']), T(g['si_logits'])
  close(o_stepfun.invert_cdf(T(g['ic_u']), t, logits), g['ic_t'], msg='invert_cdf')
  close(o_stepfun.invert_cdf(T(g['ic_u']), t, logits, use_gpu_resampling=True), g['ic_t_gpu'],
        msg='invert_cdf gather')
  for ns in [8, 32]:
    close(o_stepfun.sample(None, t, logits, ns), g[f's_det_{ns}'], msg='sample det')
    close(o_stepfun.sample(None, t, logits, ns, deterministic_center=True), g[f's_detc_{ns}'],
          msg='sample det center')
    close(o_stepfun.sample_intervals(None, t, logits, ns, domain=(0.0, 1.0)), g[f'si_det_{ns}'],
          msg='sample_intervals det')
    for sj in [1, 0]:
      out = o_stepfun.sample_intervals(T(g[f'si_jit_{ns}_{sj}_in']), t, logits, ns,
                                       single_jitter=bool(sj), domain=(0.0, 1.0))
      close(out, g[f'si_jit_{ns}_{sj}'], msg=f'sample_intervals jitter {ns} {sj}')
  close(o_stepfun.lossfun_distortion(T(g['dl_t']), T(g['dl_w'])), g['dl_loss'], msg='distortion')
  close(o_stepfun.weighted_percentile(T(g['wp_t']), T(g['wp_w']), [5, 50, 95]), g['wp'],
        rtol=1e-5, msg='weighted_percentile')


def test_sample_intervals_single_interval_known_answer():
  # reference tests/stepfun_test.py:579-586
  out = o_stepfun.sample_intervals(None, torch.tensor([3.0, 4.0]), torch.tensor([0.0]), 10)
  close(out, np.linspace(3, 4, 11), atol=1e-5, msg='linspace(3,4,11)')
  close(out, golden('stepfun')['si_single'], atol=1e-6)
  with pytest.raises(ValueError):
    o_stepfun.sample_intervals(None, torch.tensor([3.0, 4.0]), torch.tensor([0.0]), 1)


def test_searchsorted_out_of_bounds_known_answer():
  # reference tests/stepfun_test.py:79-106: queries left/right of the range collapse
  a = torch.sort(torch.rand(4, 10), dim=-1)[0]
  lo, hi = o_stepfun.searchsorted(a, a[:, :1] - 1.0)
  assert (lo == 0).all() and (hi == 0).all()
  lo, hi = o_stepfun.searchsorted(a, a[:, -1:] + 1.0)
  assert (lo == 9).all() and (hi == 9).all()


def test_distortion_loss_linear_form_matches():
  # O(S) identity used by the CUDA kernel, checked in fp64.
  g = golden('stepfun')
  t, w = T(g['dl_t']).double(), T(g['dl_w']).double()
  m = (t[..., 1:] + t[..., :-1]) / 2
  W = torch.cumsum(w, -1) - w
  M = torch.cumsum(w * m, -1) - w * m
  fast = 2 * (w * (m * W - M)).sum(-1) + (w ** 2 * (t[..., 1:] - t[..., :-1])).sum(-1) / 3
  close(fast, o_stepfun.lossfun_distortion(t, w), atol=1e-14, rtol=1e-12)


def test_render_golden():
  g = golden('render')
  o, d, radii, tdist = T(g['o']), T(g['d']), T(g['radii']), T(g['tdist'])
  for shape in ['cone', 'cylinder']:
    for diag in [0, 1]:
      m, c = o_render.cast_rays(tdist, o, d, radii, shape, diag=bool(diag))
      close(m, g[f'cast_{shape}_{diag}_mean'], msg=f'cast mean {shape} {diag}')
      close(c, g[f'cast_{shape}_{diag}_cov'], atol=1e-7, rtol=2e-4, msg=f'cast cov {shape} {diag}')
  with pytest.raises(ValueError):
    o_render.cast_rays(tdist, o, d, radii, 'sphere')
  density = T(g['density'])
  for ob in [0, 1]:
    w, a, tr = o_render.compute_alpha_weights(density, tdist, d, opaque_background=bool(ob))
    close(w, g[f'aw_{ob}_w'], atol=1e-6, msg='weights')
    close(a, g[f'aw_{ob}_alpha'], atol=1e-6, msg='alpha')
    close(tr, g[f'aw_{ob}_trans'], atol=1e-6, msg='trans')
    r = o_render.volumetric_rendering(T(g['rgbs']), T(g[f'aw_{ob}_w']), tdist, 1.0, T(g['far']), True,
                                      extras={'normals': T(g['normals']), 'roughness': T(g['rough']),
                                              'normals_pred': None})
    for k, v in r.items():
      close(v, g[f'vr_{ob}_{k}'], atol=1e-5, rtol=1e-5, msg=f'vr {ob} {k}')
  r = o_render.volumetric_rendering(T(g['rgbs']), T(g['aw_0_w']), tdist, T(g['bg']), T(g['far']), False)
  close(r['rgb'], g['vr_bg_rgb'], msg='vr bg')


def test_alpha_weights_delta_density_known_answer():
  # reference tests/render_test.py:443-463: a delta density gives one-hot weights
  tdist = torch.linspace(0.0, 1.0, 11)[None]
  density = torch.zeros(1, 10)
  density[0, 4] = 1e10
  w = o_render.compute_alpha_weights(density, tdist, torch.tensor([[0.0, 0.0, 1.0]]))[0]
  close(w, torch.nn.functional.one_hot(torch.tensor([4]), 10).float(), atol=1e-5)


def test_coord_golden():
  g = golden('coord')
  x, cov = T(g['x']), T(g['cov'])
  close(o_coord.contract(x), g['contract'], msg='contract')
  zm, zc = o_coord.track_linearize_contract(x, cov)
  close(zm, g['tl_mean'], msg='tl mean')
  close(zc, g['tl_cov'], atol=2e-7, rtol=1e-4, msg='tl cov (golden J is fp64 central differences)')
  am, ac = o_coord.track_linearize_autograd(o_coord.contract, x.double(), cov.double())
  zm64, zc64 = o_coord.track_linearize_contract(x.double(), cov.double())
  close(zc64, ac, atol=1e-12, rtol=1e-9, msg='closed-form Jacobian vs autograd')
  s = T(g['s'])
  for name, fn, near, far in [('none', None, 2.0, 6.0), ('reciprocal', 'reciprocal', 0.2, 1e6),
                              ('piecewise', 'piecewise', 0.0, 50.0), ('log', 'log', 0.5, 100.0)]:
    tn, tf = torch.full((8, 1), near), torch.full((8, 1), far)
    t_to_s, s_to_t = o_coord.construct_ray_warps(fn, tn, tf)
    tt = s_to_t(s)
    close(tt, g[f'warp_{name}_t'], rtol=1e-5, msg=f'warp {name}')
    close(t_to_s(tt), g[f'warp_{name}_s'], atol=1e-5, msg=f'warp inv {name}')
  for tag, shape, sub, mind, maxd in [('ico', 'icosahedron', 2, 0, 12), ('oct', 'octahedron', 1, 0, 16)]:
    basis = torch.tensor(geopoly.generate_basis(shape, sub), dtype=torch.float32)
    lm, lv = o_coord.lift_and_diagonalize(T(g['tl_mean']), T(g['tl_cov']), basis.T.contiguous())
    close(lm, g[f'lift_{tag}_mean'], atol=1e-6, msg='lift mean')
    close(lv, g[f'lift_{tag}_var'], atol=1e-7, rtol=1e-4, msg='lift var')
    enc = o_coord.integrated_pos_enc(T(g[f'lift_{tag}_mean']), T(g[f'lift_{tag}_var']), mind, maxd)
    close(enc, g[f'ipe_{tag}'], atol=2e-6, rtol=0, msg='ipe')
  dirs = T(g['dirs'])
  close(o_coord.pos_enc(dirs, 0, 4), g['pos_enc'], atol=1e-6, msg='pos_enc')
  close(o_coord.reflect(dirs, T(g['nrm'])), g['reflect'], atol=1e-6, msg='reflect')
  close(o_coord.l2_normalize(x[:32]), g['l2n'], atol=1e-6, msg='l2n')
  close(o_coord.generate_ide_fn(5)(dirs, T(g['kinv'])), g['ide5'], atol=2e-5, msg='ide5')
  close(o_coord.generate_ide_fn(4)(dirs, T(g['kinv'])), g['ide4'], atol=2e-5, msg='ide4')
  with pytest.raises(ValueError):
    o_coord.generate_ide_fn(6)


def test_contract_known_answers():
  # reference tests/coord_test.py:71-91: |contract(x)| <= 2, identity inside the unit ball
  x = torch.randn(1000, 3) * 10
  assert (o_coord.contract(x).norm(dim=-1) <= 2 + 1e-5).all()
  x = torch.randn(1000, 3)
  x = x / x.norm(dim=-1, keepdim=True) * torch.rand(1000, 1)
  close(o_coord.contract(x), x, atol=0, rtol=0)


def test_ide_matches_scipy_sph_harm():
  # reference tests/ref_utils_test.py:61-83 (scipy>=1.15 spells it sph_harm_y(l, m, polar, az))
  from scipy import special
  rng = np.random.default_rng(0)
  xyz = rng.normal(size=(50, 3))
  xyz /= np.linalg.norm(xyz, axis=-1, keepdims=True)
  ide = o_coord.generate_ide_fn(5)(torch.tensor(xyz, dtype=torch.float32), torch.zeros(50, 1)).numpy()
  ml, _ = o_coord.ide_tables(5)
  polar = np.arccos(xyz[:, 2])
  az = np.arctan2(xyz[:, 1], xyz[:, 0])
  expect = np.stack([special.sph_harm_y(int(l), int(m), polar, az) for m, l in ml.T], -1)
  np.testing.assert_allclose(ide[:, :ml.shape[1]], expect.real, atol=0.02)
  np.testing.assert_allclose(ide[:, ml.shape[1]:], expect.imag, atol=0.02)
