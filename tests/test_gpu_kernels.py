"""Parity of each CUDA entry point (through the C ABI) against the CPU oracle.  Needs a B200.

Tolerances: integer indices bit-exact given the same CDF; fp32 stages atol=rtol=1e-5 unless a
comment says why not; bf16 tensor-core GEMMs are compared against the same product evaluated
in fp32 from the bf16-rounded operands (error = accumulation order + one output rounding).
"""
import math

import numpy as np
import pytest
import torch

from oracle import o_coord, o_math, o_render, o_stepfun
from util import close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
  from multinerf_b200 import lib, ops as _ops
  lib.require_device()
  return _ops


def _stepfun(rng, b, n, dup=True):
  t = np.sort(rng.uniform(0, 1, (b, n + 1)).astype(np.float32), -1)
  t[:, 0], t[:, -1] = 0, 1
  if dup:
    t[::3, n // 2] = t[::3, n // 2 - 1]
  w = rng.uniform(0, 1, (b, n)).astype(np.float32) ** 3
  w /= w.sum(-1, keepdims=True)
  return torch.tensor(t), torch.tensor(w.astype(np.float32))


@pytest.mark.parametrize('P,S,dil,single', [(64, 64, 0.0103125, True), (64, 32, 0.0026220703125, True),
                                            (1, 64, 0.0, True), (128, 128, 0.0, False),
                                            (37, 17, 0.02, False)])
def test_sample_level_vs_oracle(ops, P, S, dil, single):
  rng = np.random.default_rng(P * 1000 + S)
  B = 257
  t, w = _stepfun(rng, B, P, dup=P > 4)
  use_dil = dil > 0
  anneal, pad = 0.9091, 0.0 if use_dil else 0.01
  jit = torch.tensor(rng.uniform(0, 1, (B, 1 if single else S)).astype(np.float32))
  # oracle
  if use_dil:
    td, wd = o_stepfun.max_dilate_weights(t, w, dil, domain=(0.0, 1.0), renormalize=True)
    td, wd = td[..., 1:-1], wd[..., 1:-1]
  else:
    td, wd = t, w
  logits = torch.where(td[..., 1:] > td[..., :-1], anneal * torch.log(wd + pad), torch.tensor(-math.inf))
  sd_o, idx_o, cw_o = o_stepfun.sample_intervals(jit, td, logits, S, single_jitter=single,
                                                domain=(0.0, 1.0), return_index=True)
  # device, end to end
  sd, dbg = ops.sample_level(t.cuda(), w.cuda(), S, dilation=dil, use_dilation=use_dil, anneal=anneal,
                             resample_padding=pad, jitter=(jit[:, 0] if single else jit).contiguous().cuda(),
                             single_jitter=single, want_index=True, want_debug=True)
  if use_dil:
    np.testing.assert_array_equal(dbg['tdil'].cpu().numpy(), td.numpy())      # merge == sort: exact
    close(dbg['wdil'], wd, atol=1e-7, rtol=1e-5, msg='dilated weights')
  close(dbg['cw'], cw_o, atol=2e-6, rtol=0, msg='cdf')
  close(sd, sd_o, atol=1e-5, rtol=1e-5, msg='sdist end-to-end')
  mism = (dbg['idx'].cpu().long() != idx_o)
  # end to end the CDFs differ by rounding, so an index may flip only where u sits on a knot
  assert mism.float().mean() < 2e-3, mism.float().mean()
  # integer contract: same CDF in -> identical interval indices and sdist to 1 ulp
  sd2, dbg2 = ops.sample_level(t.cuda(), w.cuda(), S, dilation=dil, use_dilation=use_dil, anneal=anneal,
                               resample_padding=pad, jitter=(jit[:, 0] if single else jit).contiguous().cuda(),
                               single_jitter=single, cw_in=cw_o.contiguous().cuda(), want_index=True)
  np.testing.assert_array_equal(dbg2['idx'].cpu().numpy(), idx_o.numpy().astype(np.int32))
  close(sd2, sd_o, atol=2e-7, rtol=1e-6, msg='sdist with shared cdf')


def test_sample_level_deterministic_and_errors(ops):
  t = torch.tensor([[3.0, 4.0]] * 5).cuda()
  w = torch.ones(5, 1).cuda()
  sd = ops.sample_level(t, w, 10, domain=(-math.inf, math.inf))
  close(sd, np.tile(np.linspace(3, 4, 11, dtype=np.float32), (5, 1)), atol=1e-5)   # stepfun_test.py:579-586
  with pytest.raises(ValueError):
    ops.sample_level(t, w, 1)


def _rays(rng, b, unit_cube=True):
  o = rng.uniform(-1, 1, (b, 3)).astype(np.float32)
  d = rng.normal(size=(b, 3)).astype(np.float32)
  d = (d / np.linalg.norm(d, axis=-1, keepdims=True) * rng.uniform(0.8, 1.2, (b, 1))).astype(np.float32)
  radii = rng.uniform(5e-4, 1e-3, (b, 1)).astype(np.float32)
  return torch.tensor(o), torch.tensor(d), torch.tensor(radii)


@pytest.mark.parametrize('name,shape,sub,maxdeg,raydist,near,far,contract,rshape', [
    ('360', 'icosahedron', 2, 12, 'reciprocal', 0.2, 1e6, True, 'cone'),
    ('blender', 'octahedron', 1, 16, None, 2.0, 6.0, False, 'cone'),
    ('llff', 'octahedron', 1, 16, None, 0.0, 1.0, False, 'cylinder'),
])
def test_encode_vs_oracle(ops, name, shape, sub, maxdeg, raydist, near, far, contract, rshape):
  from multinerf_b200 import geopoly
  rng = np.random.default_rng(7)
  B, S = 96, 32
  o, d, radii = _rays(rng, B)
  sdist = torch.tensor(np.sort(rng.uniform(0, 1, (B, S + 1)).astype(np.float32), -1))
  sdist[:, 0], sdist[:, -1] = 0, 1
  nearv, farv = torch.full((B, 1), near), torch.full((B, 1), far)
  basis = torch.tensor(geopoly.generate_basis(shape, sub), dtype=torch.float32)
  _, s_to_t = o_coord.construct_ray_warps(raydist, nearv, farv)
  tdist_o = s_to_t(sdist)
  means, covs = o_render.cast_rays(tdist_o, o, d, radii, rshape, diag=False)
  if contract:
    means, covs = o_coord.track_linearize_contract(means, covs)
  lm, lv = o_coord.lift_and_diagonalize(means, covs, basis.T.contiguous())
  enc_o = o_coord.integrated_pos_enc(lm, lv, 0, maxdeg)
  feat, f32, tdist = ops.encode(sdist.cuda(), o.cuda(), d.cuda(), radii[:, 0].contiguous().cuda(),
                                nearv[:, 0].contiguous().cuda(), farv[:, 0].contiguous().cuda(), basis.cuda(),
                                min_deg=0, max_deg=maxdeg, raydist_fn=raydist, ray_shape=rshape,
                                warp_contract=contract, want_f32=True, want_tdist=True)
  close(tdist, tdist_o, atol=0, rtol=2e-6, msg='tdist')
  F = enc_o.shape[-1]
  # (1) Bulk: 1e-5 (plus the argument-scale term: degree l multiplies the lifted mean by 2^l, so a
  #     1-ulp difference in the mean moves sin() by |x| 2^-23).
  # (2) Tail: with contraction, J cov J^T cancels ~1e11-sized terms for far samples, so the lifted
  #     variance -- in the reference's own fp32 formula -- carries a large relative error and
  #     exp(-v/2) is ill-conditioned where v ~ 1.  There the fp32 oracle itself is off from an fp64
  #     evaluation; the kernel must be no worse than a small multiple of that.
  scale = 2.0 ** (maxdeg - 1) * float(lm.abs().max()) * 2 ** -23
  tol = max(1e-5, 4 * scale)
  got = f32.view(B, S, F).cpu()
  bad = ((got - enc_o).abs() > tol).float().mean()
  assert float(bad) < 1e-3, ('ipe fp32 bulk', float(bad))
  m64, c64 = o_render.cast_rays(s_to_t(sdist).double(), o.double(), d.double(), radii.double(), rshape, diag=False)
  if contract:
    m64, c64 = o_coord.track_linearize_contract(m64, c64)
  lm64, lv64 = o_coord.lift_and_diagonalize(m64, c64, basis.double().T.contiguous())
  enc64 = o_coord.integrated_pos_enc(lm64.float(), lv64.float(), 0, maxdeg).double()
  e_gpu, e_o32 = (got.double() - enc64).abs().max(), (enc_o.double() - enc64).abs().max()
  assert float(e_gpu) <= 4 * float(e_o32) + 10 * tol, ('ipe fp32 tail', float(e_gpu), float(e_o32))
  fb = feat.float().view(B, S, -1)
  assert float(((fb[..., :F].cpu() - enc_o.to(torch.bfloat16).float()).abs() > max(8e-3, tol)).float().mean()) < 1e-3
  assert (fb[..., F:] == 0).all()
  with pytest.raises(ValueError):
    ops.encode(sdist.cuda(), o.cuda(), d.cuda(), radii[:, 0].contiguous().cuda(), nearv[:, 0].contiguous().cuda(),
               farv[:, 0].contiguous().cuda(), basis.cuda(), min_deg=0, max_deg=4, ray_shape='sphere')


def test_viewdir_enc(ops):
  rng = np.random.default_rng(3)
  B, S = 33, 5
  v = rng.normal(size=(B, 3)).astype(np.float32)
  v /= np.linalg.norm(v, axis=-1, keepdims=True)
  out = torch.full((B * S, 320), 7.0, dtype=torch.bfloat16, device='cuda')
  ops.viewdir_enc(torch.tensor(v).cuda(), S, 4, out, 256, 320)
  enc = o_coord.pos_enc(torch.tensor(v), 0, 4)
  got = out.float().cpu().view(B, S, 320)
  close(got[:, :, 256:283], enc[:, None, :].expand(B, S, 27).to(torch.bfloat16).float(), atol=8e-3)
  assert (got[:, :, 283:] == 0).all() and (got[:, :, :256] == 7).all()


CFG = dict(raydist_fn='reciprocal', opaque_background=True, density_bias=-1.0, density_noise=0.0,
           rgb_activation='sigmoid', rgb_premultiplier=1.0, rgb_bias=0.0, rgb_padding=0.001, bg_const=1.0)


def _oracle_composite(raw_d, raw_rgb, sdist, d, near, far, cfg, extras=False):
  _, s_to_t = o_coord.construct_ray_warps(cfg['raydist_fn'], near, far)
  tdist = s_to_t(sdist)
  density = torch.nn.functional.softplus(raw_d + cfg['density_bias'])
  if raw_rgb is None:
    rgb = torch.zeros(raw_d.shape + (3,))
  else:
    z = cfg['rgb_premultiplier'] * raw_rgb + cfg['rgb_bias']
    act = torch.sigmoid(z) if cfg['rgb_activation'] == 'sigmoid' else o_math.safe_exp(z)
    rgb = act * (1 + 2 * cfg['rgb_padding']) - cfg['rgb_padding']
  w = o_render.compute_alpha_weights(density, tdist, d, opaque_background=cfg['opaque_background'])[0]
  r = o_render.volumetric_rendering(rgb, w, tdist, cfg['bg_const'], far, extras)
  return w, r, density, rgb


@pytest.mark.parametrize('S,opaque,raydist,near,far,act', [
    (32, True, 'reciprocal', 0.2, 1e6, 'sigmoid'), (64, True, 'reciprocal', 0.2, 1e6, 'sigmoid'),
    (128, False, None, 2.0, 6.0, 'sigmoid'), (48, False, None, 0.0, 1.0, 'safe_exp')])
def test_composite_fwd_vs_oracle(ops, S, opaque, raydist, near, far, act):
  rng = np.random.default_rng(S)
  B = 130
  cfg = dict(CFG, raydist_fn=raydist, opaque_background=opaque, rgb_activation=act,
             rgb_bias=-5.0 if act == 'safe_exp' else 0.0, rgb_padding=0.0 if act == 'safe_exp' else 0.001)
  _, d, _ = _rays(rng, B)
  sdist = torch.tensor(np.sort(rng.uniform(0, 1, (B, S + 1)).astype(np.float32), -1))
  sdist[:, 0], sdist[:, -1] = 0, 1
  raw_d = torch.tensor(rng.normal(size=(B, S)).astype(np.float32) * 3)
  raw_d[0] = -50
  raw_d[1] = 30
  raw_rgb = torch.tensor(rng.normal(size=(B, S, 3)).astype(np.float32))
  nearv, farv = torch.full((B, 1), near), torch.full((B, 1), far)
  w_o, r_o, dens_o, rgb_o = _oracle_composite(raw_d, raw_rgb, sdist, d, nearv, farv, cfg, extras=True)
  out = ops.composite_fwd(raw_d.cuda(), raw_rgb.cuda(), sdist.cuda(), d.cuda(), nearv[:, 0].contiguous().cuda(),
                          farv[:, 0].contiguous().cuda(), cfg=cfg, want_samples=True, want_extras=True)
  close(out['density'], dens_o, msg='density')
  close(out['rgb_samples'], rgb_o, msg='rgb samples')
  close(out['weights'], w_o, atol=1e-6, rtol=1e-5, msg='weights')
  close(out['rgb'], r_o['rgb'], msg='pixel')
  close(out['acc'], r_o['acc'], msg='acc')
  dist = out['dist'].cpu()
  close(dist[:, 0], r_o['distance_mean'], rtol=2e-5, atol=1e-5, msg='distance_mean')
  # percentiles are piecewise-linear in the CDF; a knot within rounding of p moves the answer by a
  # whole interval, so compare through the CDF (|cdf(t_gpu) - p| small) on top of the bulk check
  for i, k in enumerate(['distance_percentile_5', 'distance_median', 'distance_percentile_95']):
    ref = r_o[k]
    rel = ((dist[:, 1 + i] - ref).abs() / ref.abs().clamp(min=1e-6))
    assert (rel < 1e-3).float().mean() > 0.97, (k, rel.max())
  # PropMLP levels: rgb = 0
  w_o2, r_o2, _, _ = _oracle_composite(raw_d, None, sdist, d, nearv, farv, cfg)
  out2 = ops.composite_fwd(raw_d.cuda(), None, sdist.cuda(), d.cuda(), nearv[:, 0].contiguous().cuda(),
                           farv[:, 0].contiguous().cuda(), cfg=cfg)
  close(out2['rgb'], r_o2['rgb'], msg='prop pixel')


@pytest.mark.parametrize('level,loss_type,S', [('fine', 'charb', 32), ('prop', 'mse', 64),
                                              ('fine', 'mse', 128), ('fine', 'rawnerf', 32)])
def test_composite_bwd_vs_oracle_autograd(ops, level, loss_type, S):
  from oracle import o_train
  rng = np.random.default_rng(11)
  B, Sf = 70, 32
  cfg = dict(CFG)
  _, d, _ = _rays(rng, B)

  def mk_sdist(n):
    s = torch.tensor(np.sort(rng.uniform(0, 1, (B, n + 1)).astype(np.float32), -1))
    s[:, 0], s[:, -1] = 0, 1
    return s
  sdist = mk_sdist(S)
  raw_d = torch.tensor(rng.normal(size=(B, S)).astype(np.float32) * 2, requires_grad=True)
  raw_rgb = None if level == 'prop' else torch.tensor(rng.normal(size=(B, S, 3)).astype(np.float32),
                                                     requires_grad=True)
  nearv, farv = torch.full((B, 1), 0.2), torch.full((B, 1), 1e6)
  target = torch.tensor(rng.uniform(0, 1, (B, 3)).astype(np.float32))
  lossmult = torch.tensor(rng.integers(0, 2, (B, 3)).astype(np.float32)) if loss_type == 'rawnerf' \
      else torch.ones(B, 1)
  sdist_f = mk_sdist(Sf)
  w_f = torch.tensor(rng.uniform(0, 1, (B, Sf)).astype(np.float32))
  w_f = w_f / w_f.sum(-1, keepdim=True) * 0.9

  class Cfg:
    data_loss_type = loss_type
    charb_padding = 0.001
    disable_multiscale_loss = False
    data_coarse_loss_mult = 0.1
    data_loss_mult = 1.0
    interlevel_loss_mult = 1.0
    distortion_loss_mult = 0.01
  w_o, r_o, _, _ = _oracle_composite(raw_d, raw_rgb, sdist, d, nearv, farv, cfg)
  lm = lossmult.expand(B, 3)
  data, st = o_train.compute_data_loss(target, [r_o], lm, Cfg)   # single level -> data_loss_mult
  data_mult = 1.0
  if level == 'prop':
    data = data * 0.1
    data_mult = 0.1
    hist = [dict(sdist=sdist, weights=w_o), dict(sdist=sdist_f, weights=w_f)]
    extra = o_train.interlevel_loss(hist, Cfg)
    dist_mult, inter_mult = 0.0, 1.0
  else:
    extra = o_train.distortion_loss([dict(sdist=sdist, weights=w_o)], Cfg)
    dist_mult, inter_mult = 0.01, 0.0
  loss = data + extra
  grads = torch.autograd.grad(loss, [raw_d] + ([raw_rgb] if raw_rgb is not None else []))
  stats = torch.zeros(8, device='cuda')
  inv_denom = (1.0 / lm.sum()).reshape(1).cuda()
  g_d, g_rgb = ops.composite_bwd(
      raw_d.detach().cuda(), None if raw_rgb is None else raw_rgb.detach().cuda(), sdist.cuda(), d.cuda(),
      nearv[:, 0].contiguous().cuda(), farv[:, 0].contiguous().cuda(), target.cuda(), lossmult.contiguous().cuda(),
      inv_denom, stats, cfg=cfg, loss_type=loss_type, charb_padding=0.001, data_mult=data_mult,
      distortion_mult=dist_mult, interlevel_mult=inter_mult,
      sdist_fine=sdist_f.cuda() if level == 'prop' else None,
      weights_fine=w_f.cuda() if level == 'prop' else None)
  scale = float(grads[0].abs().max())
  close(g_d, grads[0], atol=2e-5 * scale, rtol=2e-4, msg='d raw_density')
  if raw_rgb is not None:
    close(g_rgb, grads[1], atol=2e-5 * float(grads[1].abs().max()), rtol=2e-4, msg='d raw_rgb')
  st_gpu = stats.cpu()
  close(st_gpu[0], data.detach(), rtol=1e-4, atol=1e-7, msg='data loss')
  close(st_gpu[1], st['mses'][0].detach(), rtol=1e-4, atol=1e-7, msg='mse')
  close(st_gpu[2] + st_gpu[3], extra.detach(), rtol=1e-4, atol=1e-7, msg='regulariser')


def _bf(x):
  return torch.tensor(x).to(torch.bfloat16)


@pytest.mark.parametrize('impl', [1, 0])
@pytest.mark.parametrize('M,N,K', [(128, 256, 64), (256, 256, 512), (1000, 128, 320), (384, 1024, 1536),
                                   (130, 64, 128), (4096, 256, 256), (38000, 256, 64), (10000, 768, 128),
                                   # whole 256-row units: CTA-pair (cta_group::2) variant, several tiles per pair
                                   (38144, 256, 64), (10240, 1024, 256), (512, 512, 128),
                                   # CTA pairs at the 360.gin NerfMLP layer shapes (K = 1024 and the skip layer's 1536)
                                   (512, 1024, 1024), (512, 1024, 1536), (16384, 1024, 512)])
def test_gemm_fwd(ops, impl, M, N, K):
  from multinerf_b200 import lib as L
  rng = np.random.default_rng(M + N + K)
  a = _bf(rng.normal(size=(M, K)).astype(np.float32))
  w = _bf(rng.normal(size=(N, K)).astype(np.float32) / math.sqrt(K))
  bias = torch.tensor(rng.normal(size=(N,)).astype(np.float32))
  ref = torch.relu(a.float() @ w.float().T + bias)
  out = torch.full((M, N + 64), -3.0, dtype=torch.bfloat16, device='cuda')   # strided output view
  bits = torch.full((M, N // 32), -1, dtype=torch.int32, device='cuda')
  ops.gemm(L.GEMM_FWD, a.cuda(), w.cuda(), out[:, :N], m=M, n=N, k=K, act=L.ACT_RELU, bias=bias.cuda(),
           maskbits=bits, impl=impl)
  torch.cuda.synchronize()
  close(out[:, :N].float(), ref.to(torch.bfloat16).float(), atol=2e-2, rtol=1.6e-2, msg=f'fwd impl={impl}')
  assert (out[:, N:] == -3).all()
  # 1-bit ReLU mask == (stored activation > 0), bit j of word w <-> column 32w + j
  got = ((bits.cpu().long()[:, :, None] >> torch.arange(32)) & 1).reshape(M, N).bool()
  assert torch.equal(got, out[:, :N].float().cpu() > 0)
  # no activation, no bias, no mask output
  out2 = torch.empty(M, N, dtype=torch.bfloat16, device='cuda')
  ops.gemm(L.GEMM_FWD, a.cuda(), w.cuda(), out2, m=M, n=N, k=K, impl=impl)
  close(out2.float(), (a.float() @ w.float().T).to(torch.bfloat16).float(), atol=2e-2, rtol=1.6e-2, msg='fwd plain')


@pytest.mark.parametrize('impl', [1, 0])
@pytest.mark.parametrize('M,N,K', [(256, 256, 256), (1000, 1024, 1024), (384, 256, 128),
                                   # several tiles per persistent CTA: row inputs prefetched a tile ahead,
                                   # column sums resident in registers (N=768: 3 column blocks, not resident)
                                   (38000, 256, 256), (10000, 1024, 256), (10000, 768, 128),
                                   # CTA-pair variant (M % 256 == 0)
                                   (37888, 256, 256), (10240, 1024, 256), (10240, 768, 128),
                                   # CTA pairs with the long reductions of the 360.gin NerfMLP (K = 1024 / 1536)
                                   (512, 1024, 1024), (512, 1024, 1536), (16384, 1024, 1024)])
def test_gemm_dgrad(ops, impl, M, N, K):
  from multinerf_b200 import lib as L
  rng = np.random.default_rng(M + 7 * N + K)
  dy = _bf(rng.normal(size=(M, K)).astype(np.float32))           # reduction over the layer's outputs
  w_kn = _bf(rng.normal(size=(N, K)).astype(np.float32) / math.sqrt(K))   # [in, out]
  mask = _bf(rng.normal(size=(M, N)).astype(np.float32))
  rowv = torch.tensor(rng.normal(size=(M,)).astype(np.float32))
  colv = torch.tensor(rng.normal(size=(N,)).astype(np.float32))
  ref = (dy.float() @ w_kn.float().T + rowv[:, None] * colv[None, :]) * (mask.float() > 0)
  out = torch.empty(M, N, dtype=torch.bfloat16, device='cuda')
  ops.gemm(L.GEMM_DGRAD, dy.cuda(), w_kn.cuda(), out, m=M, n=N, k=K, rowv=rowv.cuda(), colv=colv.cuda(),
           mask=mask.cuda(), impl=impl)
  torch.cuda.synchronize()
  close(out.float(), ref.to(torch.bfloat16).float(), atol=3e-2, rtol=1.6e-2, msg=f'dgrad impl={impl}')
  # same mask as packed bits
  mb = (mask.float() > 0).reshape(M, N // 32, 32).long()
  words = (mb << torch.arange(32)).sum(-1)
  words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
  out2 = torch.empty(M, N, dtype=torch.bfloat16, device='cuda')
  cs = torch.full((N,), 3.0, device='cuda')
  ops.gemm(L.GEMM_DGRAD, dy.cuda(), w_kn.cuda(), out2, m=M, n=N, k=K, rowv=rowv.cuda(), colv=colv.cuda(),
           maskbits=words.cuda(), colsum=cs, impl=impl)
  torch.cuda.synchronize()
  close(out2.float(), ref.to(torch.bfloat16).float(), atol=3e-2, rtol=1.6e-2, msg=f'dgrad bits impl={impl}')
  # fused bias gradient = column sums of the (fp32, pre-rounding) output, accumulated into cs
  close(cs, ref.sum(0) + 3.0, atol=2e-2 * math.sqrt(M), rtol=2e-3, msg=f'dgrad colsum impl={impl}')
  out3 = torch.empty(M, N, dtype=torch.bfloat16, device='cuda')
  ops.gemm(L.GEMM_DGRAD, dy.cuda(), w_kn.cuda(), out3, m=M, n=N, k=K, impl=impl)
  close(out3.float(), (dy.float() @ w_kn.float().T).to(torch.bfloat16).float(), atol=3e-2, rtol=1.6e-2, msg='dgrad plain')


@pytest.mark.parametrize('impl', [1, 0])
@pytest.mark.parametrize('R,Mo,N', [(64, 128, 256), (4096, 512, 256), (8192, 320, 128), (2048, 1536, 1024),
                                    (1024, 64, 64), (65536, 256, 256),
                                    # sample counts that are not a multiple of the 64-row reduction block
                                    (2080, 128, 256), (1000, 256, 256), (37, 64, 64)])
def test_gemm_wgrad(ops, impl, R, Mo, N):
  from multinerf_b200 import lib as L
  rng = np.random.default_rng(R + Mo + N)
  x = _bf(rng.normal(size=(R, Mo)).astype(np.float32))
  dy = _bf(rng.normal(size=(R, N)).astype(np.float32))
  ref = x.float().T @ dy.float()
  out = torch.ones(Mo, N, device='cuda')          # accumulates into existing contents
  ops.gemm(L.GEMM_WGRAD, x.cuda(), dy.cuda(), out, m=Mo, n=N, k=R, impl=impl)
  torch.cuda.synchronize()
  close(out, ref + 1.0, atol=2e-3 * math.sqrt(R), rtol=1e-4, msg=f'wgrad impl={impl}')


@pytest.mark.parametrize('impl', [1, 0])
@pytest.mark.parametrize('R,Mo,N', [(65536, 256, 256), (4096, 512, 256), (8192, 1024, 1024), (16384, 1280, 256),
                                    (2080, 128, 256), (1000, 320, 128), (37, 64, 64), (5000, 256, 64)])
def test_gemm_wgrad_side_sums(ops, impl, R, Mo, N):
  """mnrf_gemm_wgrad: the weight gradient plus the bias gradient (column sums of dY) and the gradient of a
  Dense(1) head on the same activation, both taken from the operand tiles of the main loop."""
  rng = np.random.default_rng(R + 3 * Mo + N)
  x = _bf(rng.normal(size=(R, Mo)).astype(np.float32))
  dy = _bf(rng.normal(size=(R, N)).astype(np.float32))
  w = torch.tensor(rng.normal(size=(R,)).astype(np.float32))
  out = torch.ones(Mo, N, device='cuda')
  bsum = torch.full((N,), 2.0, device='cuda')
  aw = torch.full((Mo,), -1.0, device='cuda')
  ops.gemm_wgrad(x.cuda(), dy.cuda(), out, m=Mo, n=N, k=R, bsum=bsum, side_w=w.cuda(), side_aw=aw, impl=impl)
  torch.cuda.synchronize()
  close(out, x.float().T @ dy.float() + 1.0, atol=2e-3 * math.sqrt(R), rtol=1e-4, msg=f'wgrad impl={impl}')
  close(bsum, dy.float().sum(0) + 2.0, atol=2e-3 * math.sqrt(R), rtol=1e-4, msg=f'bias gradient impl={impl}')
  close(aw, (x.float() * w[:, None]).sum(0) - 1.0, atol=2e-3 * math.sqrt(R), rtol=1e-4, msg=f'head dW impl={impl}')
  # each side sum alone
  b2 = torch.zeros(N, device='cuda')
  o2 = torch.zeros(Mo, N, device='cuda')
  ops.gemm_wgrad(x.cuda(), dy.cuda(), o2, m=Mo, n=N, k=R, bsum=b2, impl=impl)
  close(b2, dy.float().sum(0), atol=2e-3 * math.sqrt(R), rtol=1e-4, msg='bias gradient alone')
  close(o2, x.float().T @ dy.float(), atol=2e-3 * math.sqrt(R), rtol=1e-4, msg='wgrad with bsum only')
  a3 = torch.zeros(Mo, device='cuda')
  o3 = torch.zeros(Mo, N, device='cuda')
  ops.gemm_wgrad(x.cuda(), dy.cuda(), o3, m=Mo, n=N, k=R, side_w=w.cuda(), side_aw=a3, impl=impl)
  close(a3, (x.float() * w[:, None]).sum(0), atol=2e-3 * math.sqrt(R), rtol=1e-4, msg='head dW alone')
  # strided operands (the activation lives inside a wider buffer)
  xb = torch.zeros(R, Mo + 64, dtype=torch.bfloat16, device='cuda')
  xb[:, :Mo] = x.cuda()
  b4 = torch.zeros(N, device='cuda')
  o4 = torch.zeros(Mo, N, device='cuda')
  ops.gemm_wgrad(xb[:, :Mo], dy.cuda(), o4, m=Mo, n=N, k=R, bsum=b4, impl=impl)
  close(o4, x.float().T @ dy.float(), atol=2e-3 * math.sqrt(R), rtol=1e-4, msg='strided x')


def test_heads_and_colsum(ops):
  rng = np.random.default_rng(5)
  for M, K, n_out in [(1000, 1024, 1), (777, 128, 3), (300, 256, 4), (515, 64, 2), (4099, 256, 1), (1, 128, 3),
                      (70001, 128, 3), (333, 192, 2)]:
    x = _bf(rng.normal(size=(M, K)).astype(np.float32))
    w = _bf(rng.normal(size=(n_out, K)).astype(np.float32) / math.sqrt(K))
    b = torch.tensor(rng.normal(size=(n_out,)).astype(np.float32))
    raw = ops.head_fwd(x.cuda(), w.cuda(), b.cuda(), n_out, K)
    close(raw, x.float() @ w.float().T + b, atol=1e-4, rtol=1e-4, msg='head fwd')
    draw = torch.tensor(rng.normal(size=(M, n_out)).astype(np.float32))
    dx = torch.empty(M, K, dtype=torch.bfloat16, device='cuda')
    dw = torch.zeros(K, n_out, device='cuda')
    db = torch.zeros(n_out, device='cuda')
    dxsum = torch.ones(K, device='cuda')
    ops.head_bwd(x.cuda(), w.cuda(), draw.cuda(), n_out, K, dx=dx, relu_mask=True, dw=dw, db=db, dxsum=dxsum)
    ref_dx = (draw @ w.float()) * (x.float() > 0)
    close(dxsum, ref_dx.sum(0) + 1.0, atol=2e-3 * math.sqrt(M), rtol=1e-4, msg='head dxsum')
    close(dx.float(), ref_dx.to(torch.bfloat16).float(), atol=1e-2, rtol=1e-2, msg='head dx')
    close(dw, x.float().T @ draw, atol=2e-3 * math.sqrt(M / 1000 + 1), rtol=1e-4, msg='head dw')
    close(db, draw.sum(0), atol=1e-3 * math.sqrt(M / 1000 + 1), rtol=1e-4, msg='head db')
    # parameter gradients only (no dx): the form the Ref-NeRF heads and the chained PropMLP's tangent pass use
    dw2 = torch.zeros(K, n_out, device='cuda')
    db2 = torch.zeros(n_out, device='cuda')
    ops.head_bwd(x.cuda(), w.cuda(), draw.cuda(), n_out, K, dx=None, dw=dw2, db=db2)
    close(dw2, x.float().T @ draw, atol=2e-3 * math.sqrt(M / 1000 + 1), rtol=1e-4, msg='head dw (no dx)')
    close(db2, draw.sum(0), atol=1e-3 * math.sqrt(M / 1000 + 1), rtol=1e-4, msg='head db (no dx)')
  x = _bf(rng.normal(size=(5000, 256)).astype(np.float32))
  out = torch.zeros(256, device='cuda')
  ops.colsum(x.cuda(), 256, out)
  close(out, x.float().sum(0), atol=2e-2, rtol=1e-4, msg='colsum')


def test_pack_weights_and_adam(ops):
  from oracle import o_train
  rng = np.random.default_rng(9)
  master = torch.tensor(rng.normal(size=(320, 128)).astype(np.float32))
  w_nk = torch.empty(128, 320, dtype=torch.bfloat16, device='cuda')
  w_kn = torch.empty(320, 128, dtype=torch.bfloat16, device='cuda')
  ops.pack_weights(master.cuda(), w_nk, w_kn)
  assert torch.equal(w_kn.cpu(), master.to(torch.bfloat16))
  assert torch.equal(w_nk.cpu(), master.T.contiguous().to(torch.bfloat16))
  # all layers of a module in one launch (ragged shapes, an item without the w_kn copy)
  shapes = [(320, 128), (64, 3), (100, 257), (512, 256), (33, 1)]
  masters = [torch.tensor(rng.normal(size=sh).astype(np.float32)).cuda() for sh in shapes]
  nks = [torch.full((o, i), 7.0, dtype=torch.bfloat16, device='cuda') for i, o in shapes]
  kns = [None if j == 1 else torch.full((i, o), 7.0, dtype=torch.bfloat16, device='cuda') for j, (i, o) in enumerate(shapes)]
  table = ops.pack_table(list(zip(masters, nks, kns)), 'cuda')
  ops.pack_weights_batched(table)
  torch.cuda.synchronize()
  for mst, nk, kn in zip(masters, nks, kns):
    assert torch.equal(nk, mst.T.contiguous().to(torch.bfloat16))
    if kn is not None:
      assert torch.equal(kn, mst.to(torch.bfloat16))

  class Cfg:
    adam_beta1, adam_beta2, adam_eps = 0.9, 0.999, 1e-6
  n = 100003
  p = torch.tensor(rng.normal(size=n).astype(np.float32))
  g = torch.tensor(rng.normal(size=n).astype(np.float32) * 1e-3)
  g[5] = float('nan')
  m = torch.tensor(rng.normal(size=n).astype(np.float32) * 1e-4)
  v = torch.tensor(rng.uniform(size=n).astype(np.float32) * 1e-8)
  gc = g.clone()
  norm = torch.sqrt((gc ** 2).nansum())
  # reference semantics: norm includes NaN -> mult NaN -> nan_to_num -> 0 everywhere; test the
  # finite case for values and the NaN case for "no NaN reaches the parameters"
  gf = torch.nan_to_num(gc)
  mult = torch.clamp(1e-3 / (o_train.EPS + torch.sqrt((gf ** 2).sum())), max=1.0)
  p_ref, m_ref, v_ref = o_train.adam_update(p, mult * gf, m, v, 6, 1.5e-3, Cfg)
  pc, mc, vc = p.cuda(), m.cuda(), v.cuda()
  scratch = torch.zeros(1, device='cuda')
  ops.clip_adam(pc, gf.cuda(), mc, vc, scratch, step=7, lr=1.5e-3, beta1=0.9, beta2=0.999, eps=1e-6,
                grad_max_val=0.0, grad_max_norm=1e-3)
  close(pc, p_ref, atol=1e-7, rtol=1e-5, msg='adam p')
  close(mc, m_ref, atol=1e-9, rtol=1e-5, msg='adam m')
  close(vc, v_ref, atol=1e-14, rtol=1e-5, msg='adam v')
  pc2 = p.cuda()
  ops.clip_adam(pc2, g.cuda(), m.cuda(), v.cuda(), scratch, step=7, lr=1.5e-3, beta1=0.9, beta2=0.999,
                eps=1e-6, grad_max_val=0.0, grad_max_norm=1e-3)
  assert torch.isfinite(pc2).all()
  # Reference order (train_utils.py:200-218 then :328): value clip -> norm clip -> nan_to_num.  jnp.clip and
  # jnp.minimum propagate NaN, so one NaN makes the module's mult NaN and the WHOLE module's gradient 0:
  # Adam then sees g = 0 everywhere (moments decay, parameters move by the momentum term only).
  zero = torch.zeros(n)
  for max_val in (0.0, 0.1):
    p_z, m_z, v_z = o_train.adam_update(p, zero, m, v, 6, 1.5e-3, Cfg)
    pc3, mc3, vc3 = p.cuda(), m.cuda(), v.cuda()
    ops.clip_adam(pc3, g.cuda(), mc3, vc3, scratch, step=7, lr=1.5e-3, beta1=0.9, beta2=0.999, eps=1e-6,
                  grad_max_val=max_val, grad_max_norm=1e-3)
    close(pc3, p_z, atol=1e-7, rtol=1e-5, msg='adam p after a NaN gradient (module update zeroed)')
    close(mc3, m_z, atol=1e-9, rtol=1e-5, msg='adam m after a NaN gradient')
  # without norm clipping only the NaN element itself is zeroed (nan_to_num), also under a value clip
  g2 = g.clone()
  gf2 = torch.nan_to_num(g2).clamp(-1e-3, 1e-3)
  p_r, m_r, _ = o_train.adam_update(p, gf2, m, v, 6, 1.5e-3, Cfg)
  pc4, mc4 = p.cuda(), m.cuda()
  ops.clip_adam(pc4, g2.cuda(), mc4, v.cuda(), scratch, step=7, lr=1.5e-3, beta1=0.9, beta2=0.999, eps=1e-6,
                grad_max_val=1e-3, grad_max_norm=0.0)
  close(mc4, m_r, atol=1e-9, rtol=1e-5, msg='adam m, value clip with a NaN element')
  assert float(mc4[5]) == float(0.9 * m[5])          # NaN -> 0, not -grad_max_val


def test_composite_diffuse_specular_mode(ops):
  """rgb_mode 1 (Ref-NeRF, models.py:588-602): value and gradients vs oracle autograd."""
  from oracle import o_train
  rng = np.random.default_rng(31)
  B, S = 50, 32
  cfg = dict(CFG, raydist_fn=None, opaque_background=False, density_bias=0.5, rgb_mode=1)
  _, d, _ = _rays(rng, B)
  sdist = torch.tensor(np.sort(rng.uniform(0, 1, (B, S + 1)).astype(np.float32), -1))
  nearv, farv = torch.full((B, 1), 2.0), torch.full((B, 1), 6.0)
  leaves = [torch.tensor(rng.normal(size=sh).astype(np.float32) * sc, requires_grad=True)
            for sh, sc in [((B, S), 2.0), ((B, S, 3), 1.0), ((B, S, 3), 1.5), ((B, S, 3), 1.0)]]
  raw_d, raw_rgb, raw_dif, raw_tint = leaves
  extra = torch.tensor(rng.normal(size=(B, S)).astype(np.float32) * 1e-3)
  target = torch.tensor(rng.uniform(0, 1, (B, 3)).astype(np.float32))

  def colour(raw_rgb, raw_dif, raw_tint):
    spec = torch.sigmoid(raw_tint) * torch.sigmoid(raw_rgb)
    lin = spec + torch.sigmoid(raw_dif - math.log(3.0))
    return torch.clamp(o_math.linear_to_srgb(lin), 0.0, 1.0) * (1 + 2 * 0.001) - 0.001
  _, s_to_t = o_coord.construct_ray_warps(None, nearv, farv)
  tdist = s_to_t(sdist)
  dens = torch.nn.functional.softplus(raw_d + 0.5)
  w = o_render.compute_alpha_weights(dens, tdist, d)[0]
  c = colour(raw_rgb, raw_dif, raw_tint)
  pix = o_render.volumetric_rendering(c, w, tdist, 1.0, farv, False)['rgb']
  loss = ((pix - target) ** 2).sum() / (3 * B) + (w * extra).sum()
  grads = torch.autograd.grad(loss, leaves)
  out = ops.composite_fwd(raw_d.detach().cuda(), raw_rgb.detach().cuda(), sdist.cuda(), d.cuda(),
                          nearv[:, 0].contiguous().cuda(), farv[:, 0].contiguous().cuda(), cfg=cfg,
                          raw_diffuse=raw_dif.detach().cuda(), raw_tint=raw_tint.detach().cuda(), want_samples=True)
  close(out['rgb_samples'], c.detach(), msg='diffuse+specular colour')
  close(out['rgb'], pix.detach(), msg='pixel')
  stats = torch.zeros(8, device='cuda')
  d_dif = torch.empty(B, S, 3, device='cuda')
  d_tint = torch.empty(B, S, 3, device='cuda')
  g_d, g_rgb = ops.composite_bwd(
      raw_d.detach().cuda(), raw_rgb.detach().cuda(), sdist.cuda(), d.cuda(), nearv[:, 0].contiguous().cuda(),
      farv[:, 0].contiguous().cuda(), target.cuda(), torch.ones(B, 1).cuda(), torch.tensor([1.0 / (3 * B)]).cuda(),
      stats, cfg=cfg, loss_type='mse', charb_padding=0.001, data_mult=1.0, distortion_mult=0.0, interlevel_mult=0.0,
      raw_diffuse=raw_dif.detach().cuda(), raw_tint=raw_tint.detach().cuda(), extra_dw=extra.cuda(),
      d_raw_diffuse=d_dif, d_raw_tint=d_tint)
  for got, ref, name in [(g_d, grads[0], 'd raw_density'), (g_rgb, grads[1], 'd raw_rgb'),
                         (d_dif, grads[2], 'd raw_diffuse'), (d_tint, grads[3], 'd raw_tint')]:
    close(got, ref, atol=2e-5 * float(ref.abs().max()), rtol=3e-4, msg=name)


@pytest.mark.parametrize('mode', ['refnerf', 'viewdir_pe_normals'])
def test_refdir_stage_vs_oracle_autograd(ops, mode):
  """normals / roughness / reflection / IDE (or PE) / n.v slab, the two normal losses, and the
  adjoint of all of it (csrc/refnerf.cu) against the oracle differentiated by torch autograd."""
  from multinerf_b200 import ref_utils
  rng = np.random.default_rng(41)
  B, S = 37, 8
  M = B * S
  ide = mode == 'refnerf'
  deg = 5 if ide else 4
  v = rng.normal(size=(B, 3)).astype(np.float32)
  v /= np.linalg.norm(v, axis=-1, keepdims=True)
  vt = torch.tensor(v)
  leaves = [torch.tensor(rng.normal(size=sh).astype(np.float32) * sc, requires_grad=True)
            for sh, sc in [((B, S, 3), 1.0), ((B, S, 1), 1.0), ((3, B, S), 30.0)]]
  grad_pred, raw_rough, rgd = leaves
  w = torch.tensor(rng.uniform(0, 0.2, (B, S)).astype(np.float32))
  dim = ref_utils.ide_dim(5) if ide else 3 + 6 * deg
  ncols = dim + 1
  col0, ld = 64, 64 + 128
  g_slab = torch.tensor(rng.normal(size=(B, S, ncols)).astype(np.float32)).to(torch.bfloat16).float()
  om, pm = 0.1 / B, 3e-4 / B
  # oracle
  n_pred = -o_coord.l2_normalize(grad_pred)
  n_den = -o_coord.l2_normalize(rgd.permute(1, 2, 0))
  rough = torch.nn.functional.softplus(raw_rough - 1.0)
  if ide:
    refd = o_coord.reflect(-vt[:, None, :], n_pred)
    enc = o_coord.generate_ide_fn(5)(refd, rough)
  else:
    enc = o_coord.pos_enc(vt, 0, deg)[:, None, :].expand(B, S, dim)
  ndv = (n_pred * vt[:, None, :]).sum(-1, keepdim=True)
  slab_o = torch.cat([enc, ndv], -1)
  l_or = om * (w * torch.clamp((n_pred * -vt[:, None, :]).sum(-1), max=0.0) ** 2).sum()
  l_pn = pm * (w * (1.0 - (n_den * n_pred).sum(-1))).sum()
  loss = (slab_o * g_slab).sum() + l_or + l_pn
  grads = torch.autograd.grad(loss, leaves, allow_unused=True)
  grads = [torch.zeros_like(x) if g_ is None else g_ for g_, x in zip(grads, leaves)]   # PE ignores roughness
  # device
  m, l, mat = ref_utils.ide_tables(5)
  mat_d = torch.tensor(mat, dtype=torch.float32).cuda().contiguous()
  ml_d = torch.tensor(np.stack([m, l]), dtype=torch.int32).cuda().contiguous()
  desc = ops.refdir_desc(M, S, use_pred_normals=True, use_density_normals=True, use_reflections=ide, use_ide=ide,
                         use_n_dot_v=True, use_roughness=True, deg_view=deg, ide_n=len(m), roughness_bias=-1.0,
                         ld=ld, col0=col0, col_end=ld)
  slab = torch.full((M, ld), 5.0, dtype=torch.bfloat16, device='cuda')
  npd, nd_, rgh, edw = (torch.empty(M, 3, device='cuda'), torch.empty(M, 3, device='cuda'),
                        torch.empty(M, device='cuda'), torch.empty(M, device='cuda'))
  gp_c = grad_pred.detach().reshape(M, 3).contiguous().cuda()
  rr_c = raw_rough.detach().reshape(M).contiguous().cuda()
  rgd_c = rgd.detach().reshape(3, M).contiguous().cuda()
  ops.refdir_fwd(desc, mat_d, ml_d, gp_c, rr_c, rgd_c, vt.cuda(), npd, nd_, rgh, slab, om, pm, True, edw)
  torch.cuda.synchronize()
  close(npd.cpu().view(B, S, 3), n_pred.detach(), msg='normals_pred')
  close(nd_.cpu().view(B, S, 3), n_den.detach(), atol=1e-5, rtol=1e-4, msg='normals')
  close(rgh.cpu().view(B, S, 1), rough.detach(), msg='roughness')
  got = slab.float().cpu().view(B, S, ld)
  close(got[..., col0:col0 + ncols], slab_o.detach().to(torch.bfloat16).float(), atol=1.6e-2, rtol=1.6e-2, msg='slab')
  assert (got[..., :col0] == 5).all() and (got[..., col0 + ncols:] == 0).all()
  ref_edw = om * torch.clamp((n_pred * -vt[:, None, :]).sum(-1), max=0.0) ** 2 + pm * (1.0 - (n_den * n_pred).sum(-1))
  close(edw.cpu().view(B, S), ref_edw.detach(), atol=1e-9, rtol=1e-4, msg='extra_dw')
  # backward
  d_slab = torch.zeros(M, ld, dtype=torch.bfloat16, device='cuda')
  d_slab[:, col0:col0 + ncols] = g_slab.reshape(M, ncols).to(torch.bfloat16).cuda()
  d_gp, d_rr, d_rgd = torch.empty(M, 3, device='cuda'), torch.empty(M, device='cuda'), torch.empty(3, M, device='cuda')
  stats = torch.zeros(8, device='cuda')
  drd = torch.tensor(rng.normal(size=M).astype(np.float32)).cuda()
  ddf = torch.tensor(rng.normal(size=(M, 3)).astype(np.float32)).cuda()
  ops.refdir_bwd(desc, mat_d, ml_d, gp_c, rr_c, rgd_c, vt.cuda(), w.reshape(M).contiguous().cuda(), d_slab, om, pm,
                 True, drd, ddf, None, d_gp, d_rr, d_rgd, stats)
  torch.cuda.synchronize()
  for got_, ref, name in [(d_gp.cpu().view(B, S, 3), grads[0], 'd grad_pred'),
                          (d_rr.cpu().view(B, S, 1), grads[1], 'd raw_rough'),
                          (d_rgd.cpu().view(3, B, S), grads[2], 'd raw_grad_density')]:
    close(got_, ref, atol=2e-4 * float(ref.abs().max()) + 1e-9, rtol=2e-3, msg=name)
  close(stats[4].cpu(), l_or.detach(), rtol=1e-4, atol=1e-9, msg='orientation loss')
  close(stats[5].cpu(), l_pn.detach(), rtol=1e-4, atol=1e-9, msg='pred-normal loss')
  hs = d_slab.float().cpu()[:, col0:col0 + 11]
  close(hs[:, 0], drd.cpu().to(torch.bfloat16).float(), atol=0, rtol=0, msg='head slab density')
  close(hs[:, 1:4], d_gp.cpu().to(torch.bfloat16).float(), atol=0, rtol=0, msg='head slab grad_pred')
  close(hs[:, 4:7], ddf.cpu().to(torch.bfloat16).float(), atol=0, rtol=0, msg='head slab diffuse')
  assert (hs[:, 7:10] == 0).all()
  close(hs[:, 10], d_rr.cpu().to(torch.bfloat16).float(), atol=0, rtol=0, msg='head slab roughness')


def test_outer_mask_and_gemm_mask_mod_addend(ops):
  from multinerf_b200 import lib as L
  rng = np.random.default_rng(51)
  M, N, K = 256, 128, 192
  bits = torch.tensor(rng.integers(-2 ** 31, 2 ** 31, (M, N // 32)), dtype=torch.int32)
  maskb = ((bits.long()[:, :, None] >> torch.arange(32)) & 1).reshape(M, N).bool()
  rowv = torch.tensor(rng.normal(size=3 * M).astype(np.float32))
  colv = torch.tensor(rng.normal(size=N).astype(np.float32))
  out = torch.empty(3 * M, N, dtype=torch.bfloat16, device='cuda')
  ops.outer_mask(rowv.cuda(), colv.cuda(), bits.cuda(), out, rows=3 * M, n=N, mask_mod=M)
  ref = (rowv[:, None] * colv[None, :]) * maskb.repeat(3, 1)
  close(out.float(), ref.to(torch.bfloat16).float(), atol=0, rtol=0, msg='outer_mask')
  for impl in [1, 0]:
    a = _bf(rng.normal(size=(3 * M, K)).astype(np.float32))
    w = _bf(rng.normal(size=(N, K)).astype(np.float32) / math.sqrt(K))
    add = _bf(rng.normal(size=(3 * M, N)).astype(np.float32))
    o2 = torch.empty(3 * M, N, dtype=torch.bfloat16, device='cuda')
    ops.gemm(L.GEMM_DGRAD, a.cuda(), w.cuda(), o2, m=3 * M, n=N, k=K, maskbits=bits.cuda(), mask_mod=M,
             addend=add.cuda(), impl=impl)
    ref2 = (a.float() @ w.float().T) * maskb.repeat(3, 1) + add.float()
    close(o2.float(), ref2.to(torch.bfloat16).float(), atol=3e-2, rtol=1.6e-2, msg=f'mask_mod+addend impl={impl}')


def test_encode_tangent_features(ops):
  """d(IPE feature)/d(mean) against torch autograd of the oracle (no contraction)."""
  from multinerf_b200 import geopoly
  rng = np.random.default_rng(61)
  B, S, maxdeg = 40, 16, 16
  o, d, radii = _rays(rng, B)
  o = o * 3
  sdist = torch.tensor(np.sort(rng.uniform(0, 1, (B, S + 1)).astype(np.float32), -1))
  nearv, farv = torch.full((B, 1), 2.0), torch.full((B, 1), 6.0)
  basis = torch.tensor(geopoly.generate_basis('octahedron', 1), dtype=torch.float32)
  _, s_to_t = o_coord.construct_ray_warps(None, nearv, farv)
  means, covs = o_render.cast_rays(s_to_t(sdist), o, d, radii, 'cone', diag=False)
  means = means.detach().requires_grad_(True)
  lm, lv = o_coord.lift_and_diagonalize(means, covs, basis.T.contiguous())
  enc = o_coord.integrated_pos_enc(lm, lv, 0, maxdeg)          # [B,S,F]
  F = enc.shape[-1]
  jac = torch.stack([torch.autograd.grad(enc[..., f].sum(), means, retain_graph=True)[0] for f in range(F)], -1)
  M = B * S
  feat = torch.empty(M, 128, dtype=torch.bfloat16, device='cuda')
  tfeat = torch.empty(3 * M, 128, dtype=torch.bfloat16, device='cuda')
  ops.encode(sdist.cuda(), o.cuda(), d.cuda(), radii[:, 0].contiguous().cuda(), nearv[:, 0].contiguous().cuda(),
             farv[:, 0].contiguous().cuda(), basis.cuda(), min_deg=0, max_deg=maxdeg, feat=feat, feat_cols=128,
             tfeat=tfeat)
  got = tfeat.float().cpu().view(3, B, S, 128)[..., :F]          # [dir, B, S, F]
  ref = jac.permute(2, 0, 1, 3)                                  # [3, B, S, F]
  scale = float(ref.abs().max())
  assert float(((got - ref).abs() > 1e-2 * ref.abs() + 2e-4 * scale).float().mean()) < 2e-3
  close(feat.float().cpu().view(B, S, 128)[..., :F], enc.detach().to(torch.bfloat16).float(), atol=8e-3, rtol=0,
        msg='features unchanged')
