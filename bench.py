#!/usr/bin/env python
"""Benchmark of the hot path: train-step rays/sec @ 16384 rays x (64+64+32) samples (360.gin).

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU, NCCL)
  python bench.py --impl reference --gpus N --steps K ...  CPU arm: the oracle restatement of
                                                           the reference train step (JAX is not
                                                           installable in this image)
Prints ONE JSON line (see README/DESIGN.md for the field contract).  A "step" is one full
train step (forward 3 levels, losses, backward, grad all-reduce, clip+Adam, weight repack) on
one synthetic batch of `batch_size` rays (global; sharded B/N per GPU, as train.py:52-53).
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over
ranks; the activations written per step (>10 GB) exceed the 126 MB L2, so no flush is needed.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

# The CPU arm must own the host's threads.  torchrun exports OMP_NUM_THREADS=1 to every rank, which
# (set before the OpenMP / MKL runtimes start) would throttle the reference arm by two orders of
# magnitude, so that arm resets the variables before numpy / torch are imported.
if '--impl' in sys.argv and sys.argv[sys.argv.index('--impl') + 1:sys.argv.index('--impl') + 2] == ['reference'] \
    or '--impl=reference' in sys.argv:
  _t = str(min(os.cpu_count() or 1, int(os.environ.get('MNRF_CPU_THREADS', '32'))))
  for _k in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS'):
    os.environ[_k] = _t

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_FLOP_PER_RAY = 638435328           # 2 * MACs of Prop x128 + Nerf x32
TRAIN_FLOP_PER_RAY = 1815994368        # 3 x fwd - unneeded dgrad (canonical, never overstated)


def synth_batch(seed, B):
  """SURVEY.md 8(d) config 2 recipe: unit-cube origins, normalised directions x U(.8,1.2)."""
  rng = np.random.default_rng(seed)
  f = np.float32
  o = rng.uniform(-1, 1, (B, 3)).astype(f)
  d = rng.normal(size=(B, 3))
  d /= np.linalg.norm(d, axis=-1, keepdims=True)
  v = d.astype(f)
  d = (d * rng.uniform(0.8, 1.2, (B, 1))).astype(f)
  rays = dict(origins=o, directions=d, viewdirs=v, radii=rng.uniform(5e-4, 1e-3, (B, 1)).astype(f),
              imageplane=np.zeros((B, 2), f), lossmult=np.ones((B, 1), f),
              near=np.full((B, 1), 0.2, f), far=np.full((B, 1), 1e6, f),
              cam_idx=np.zeros((B, 1), np.int32))
  rgb = rng.uniform(0, 1, (B, 3)).astype(f)
  return rays, rgb


def peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as fh:
      p = json.load(fh)
    return p.get('bf16_tflops_sustained', p.get('bf16_tflops')), p.get('hbm_gbs'), 'measured'
  return 1400.0, 6650.0, 'fallback'      # B200_PROFILING.md fallback (sustained figure)


class ClockSampler:
  """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
       'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
       'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index):
    self.idx = gpu_index
    self.proc = None
    self.path = None

  def start(self):
    try:
      fd, self.path = tempfile.mkstemp(suffix='.csv')
      os.close(fd)
      self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                    '-lms', '100', '-i', str(self.idx)],
                                   stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
    except Exception:  # pylint: disable=broad-except
      self.proc = None

  def stop(self):
    out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
    if self.proc is None:
      return out
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except Exception:  # pylint: disable=broad-except
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    try:
      for line in open(self.path):
        p = [x.strip() for x in line.split(',')]
        if len(p) < 9:
          continue
        sm.append(float(p[1]))
        mx.append(float(p[2]))
        for name, val in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'],
                             p[5:9]):
          if val.lower().startswith('active'):
            reasons.add(name)
      os.unlink(self.path)
    except Exception:  # pylint: disable=broad-except
      pass
    if sm:
      out = {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
             'samples': len(sm)}
    return out


def run_ours(args):
  import torch
  import torch.distributed as dist
  from multinerf_b200 import configs, models, ops, train_utils, utils

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  dev = torch.device('cuda', local)
  bundle = configs.bundle_360()
  B_global = args.batch_size
  if args.scaling == 'weak':
    B_global *= world
  assert B_global % world == 0
  B = B_global // world
  model, variables = models.construct_model(2, None, bundle, device=dev)
  step_fn = train_utils.create_train_step(model, bundle.config, use_graph=not args.no_graph)
  state = train_utils.TrainState(variables)
  gen = torch.Generator(device=dev)
  gen.manual_seed(1234 + rank)

  # host batches in pinned memory; every step copies its batch H2D inside the timed region (e2e)
  nbuf = 4
  host = []
  for i in range(nbuf):
    rays_np, rgb_np = synth_batch(100 + i, B_global)
    sl = slice(rank * B, (rank + 1) * B)
    hr = {k: torch.from_numpy(np.ascontiguousarray(v[sl])).pin_memory() for k, v in rays_np.items()}
    host.append((hr, torch.from_numpy(np.ascontiguousarray(rgb_np[sl])).pin_memory()))
  h2d_bytes = sum(t.numel() * t.element_size() for t in host[0][0].values()) + host[0][1].numel() * 4

  def device_batch(i):
    hr, hrgb = host[i % nbuf]
    rays = utils.Rays(**{k: v.to(dev, non_blocking=True) for k, v in hr.items()})
    return utils.Batch(rays=model._prep_rays(rays), rgb=hrgb.to(dev, non_blocking=True))

  resident = [device_batch(i) for i in range(nbuf)]
  torch.cuda.synchronize()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(n, e2e):
    nonlocal state
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loss_host = None
    e0.record()
    for i in range(n):
      batch = device_batch(i) if e2e else resident[i % nbuf]
      state, stats, _ = step_fn(gen, state, batch, None, 0.5)
      if e2e:
        loss_host = stats._buf.to('cpu', non_blocking=False)     # D2H read of the step's losses
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
      t = torch.tensor([ms], device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = float(t)
    return ms, loss_host

  timed(args.warmup, False)
  ops.LAUNCHES = 0
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  ms, _ = timed(args.steps, False)
  clocks = sampler.stop() if rank == 0 else {}
  launches_total = ops.LAUNCHES                 # our kernels launched inside the timed region (all K steps)
  launches = launches_total // max(1, args.steps)
  ms_e2e, loss_host = timed(args.steps, True)

  # dominant kernel (tcgen05 GEMM, all three modes): CUDA events around every launch of one
  # extra step; achieved = canonical train FLOPs of the step / time spent inside the GEMMs
  eager_fn = train_utils.create_train_step(model, bundle.config, use_graph=False)
  ops.GEMM_EVENTS = []
  barrier()
  state, _, _ = eager_fn(gen, state, resident[0], None, 0.5)
  barrier()
  evs = ops.GEMM_EVENTS
  ops.GEMM_EVENTS = None
  torch.cuda.synchronize()
  gemm_ms = sum(a.elapsed_time(b) for a, b, _ in evs)
  gemm_flops = sum(f for _, _, f in evs)

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return
  rays_per_s = B_global * args.steps / (ms / 1e3)
  rays_per_s_e2e = B_global * args.steps / (ms_e2e / 1e3)
  peak_tf, peak_hbm, peak_kind = peaks()
  canon_flops_step = TRAIN_FLOP_PER_RAY * B          # per rank
  achieved = canon_flops_step / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
  # DRAM bytes per GEMM launch from the committed ncu --set full capture (tools/summarize_profile.py)
  traffic = None
  tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'gemm_tc_traffic.json')
  if os.path.exists(tpath) and B == 16384:
    with open(tpath) as f:
      traffic = json.load(f).get('dram_bytes_per_launch')
  n_gemm = max(1, len(evs))
  out = {
      'metric': 'train-step rays/sec @16384 rays x (64+64+32) samples',
      'value': rays_per_s, 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': args.scaling,
      'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
      'config': {'workload': 'mip-NeRF 360 (360.gin) train step: %d rays x (64+64+32) samples, '
                             'PropMLP 4x256, NerfMLP 8x1024' % B_global,
                 'global_batch': B_global, 'rays_per_gpu': B, 'parallelism': f'dp{world}',
                 'cuda_graphs': not args.no_graph,
                 'l2_flush': 'not needed: >10 GB of activations streamed per step (L2 = 126 MB)'},
      'e2e': {'value': rays_per_s_e2e, 'unit': 'rays/s', 'h2d_bytes_per_step': int(h2d_bytes * world),
              'd2h_bytes_per_step': int(loss_host.numel() * 4 * world) if loss_host is not None else 0,
              'ms_per_step': ms_e2e / args.steps},
      'gpu_launches': int(launches_total), 'gpu_launches_per_step': int(launches),
      'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s',
                   'frac': achieved / peak_tf if peak_tf else None, 'traffic': traffic,
                   'traffic_unit': 'DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)',
                   'launches_per_step': n_gemm, 'avg_launch_ms': gemm_ms / n_gemm,
                   'algorithmic_flop_per_launch': canon_flops_step / n_gemm,
                   'kernel': 'gemm_tc_kernel (tcgen05 fwd+dgrad+wgrad)', 'peak_kind': peak_kind + ' sustained bf16',
                   'gemm_ms_per_step': gemm_ms, 'gemm_share_of_step': gemm_ms / (ms / args.steps),
                   'executed_tflops': gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0,
                   'whole_step_frac': rays_per_s / world * TRAIN_FLOP_PER_RAY / 1e12 / peak_tf},
      'clocks': clocks,
  }
  if world == 1 and not args.no_cpu_baseline:
    out['cpu_baseline'] = cpu_baseline(args.cpu_rays, 1, 1)
  print(json.dumps(out), flush=True)
  if world > 1:
    dist.destroy_process_group()


def cpu_baseline(n_rays, steps, warmup):
  """The oracle (CPU restatement of the reference train step), all host threads, fp32."""
  import torch
  from multinerf_b200 import configs, geopoly
  from oracle import o_train
  cores = os.cpu_count() or 1
  # "all the host threads it can use": torch-CPU's intra-op pool stops scaling (and degrades) past a
  # few dozen threads on these many-small-op graphs, so use min(cores, 32) and say so.
  threads = min(cores, int(os.environ.get('MNRF_CPU_THREADS', '32')))
  torch.set_num_threads(threads)
  bundle = configs.bundle_360()
  rng = np.random.default_rng(2)
  bases = {}
  shapes = {}

  def init_tree(cfg, F):
    # layer table of internal/models.py for the non-reflective MLP (same as MLPPlan)
    tree, k, x = {}, 0, F
    W = cfg.net_width

    def dense(i, o):
      nonlocal k
      lim = math.sqrt(6.0 / i)
      tree[f'Dense_{k}'] = {'kernel': torch.tensor(rng.uniform(-lim, lim, (i, o)).astype(np.float32)),
                            'bias': torch.zeros(o)}
      k += 1
    for i in range(cfg.net_depth):
      dense(x, W)
      x = W + F if (i % cfg.skip_layer == 0 and i > 0) else W
    dense(x, 1)
    if not cfg.disable_rgb:
      dense(x, cfg.bottleneck_width)
      dense(cfg.bottleneck_width + 3 + 6 * cfg.deg_view, cfg.net_width_viewdirs)
      dense(cfg.net_width_viewdirs, 3)
    return tree
  params = {}
  for name, cfg, key in [('NerfMLP_0', bundle.nerf_mlp, 'nerf'), ('PropMLP_0', bundle.prop_mlp, 'prop')]:
    basis = geopoly.generate_basis(cfg.basis_shape, cfg.basis_subdivisions).astype(np.float32)
    bases[key] = basis
    params[name] = init_tree(cfg, 2 * basis.shape[0] * (cfg.max_deg_point - cfg.min_deg_point))
  rays_np, rgb_np = synth_batch(7, n_rays)

  class R:
    pass
  rays = R()
  for k2, v in rays_np.items():
    setattr(rays, k2, torch.tensor(v))
  rays.exposure_idx = None
  rays.exposure_values = None
  target = torch.tensor(rgb_np)
  rand = {'jitter': [torch.rand(n_rays, 1) for _ in range(3)]}
  opt = {'count': 0, 'mu': {}, 'nu': {}}
  times = []
  for i in range(warmup + steps):
    t0 = time.perf_counter()
    params, opt, _, _ = o_train.train_step(params, opt, bundle, bases, rays, target, 0.5, rand=rand)
    times.append(time.perf_counter() - t0)
  t = float(np.mean(times[warmup:]))
  return {'value': n_rays / t, 'unit': 'rays/s', 'cores': threads, 'host_cores': cores, 'kind': 'port',
          'sample': f'{n_rays} rays x (64+64+32) samples of the same 360.gin train step, fp32 torch-CPU, '
                    f'{steps} timed step(s) after {warmup} warm-up; CPU restatement of the reference '
                    '(JAX/Flax are not installable in this image)',
          's_per_step': t}


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  base = cpu_baseline(args.cpu_rays, max(1, args.steps), max(0, min(args.warmup, 3)))
  out = {
      'impl': 'reference', 'metric': 'train-step rays/sec @16384 rays x (64+64+32) samples',
      'value': base['value'], 'unit': 'rays/s', 'n_gpus': int(os.environ.get('WORLD_SIZE', '1')),
      'steps': max(1, args.steps), 'warmup': max(0, min(args.warmup, 3)),
      'ms_per_step': base['s_per_step'] * 1e3, 'higher_is_better': True, 'scaling': args.scaling,
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': 'mip-NeRF 360 (360.gin) train step, bounded sample of %d rays per step on the '
                             'host CPU (reference arm = oracle port; JAX unavailable)' % args.cpu_rays},
      'cpu_baseline': base,
      'e2e': {'value': base['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }
  print(json.dumps(out), flush=True)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--batch_size', type=int, default=16384)
  ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'])
  ap.add_argument('--cpu_rays', type=int, default=256)
  ap.add_argument('--no_cpu_baseline', action='store_true')
  ap.add_argument('--no_graph', action='store_true', help='launch every kernel from Python (no CUDA graphs)')
  args = ap.parse_args()
  if args.impl == 'reference':
    # exactly K timed steps; each step is a bounded ray sample of the workload, shrunk for large K so
    # that the whole run stays within about a minute of CPU time (256 rays ~ 0.8 s per step)
    if args.steps > 60:
      args.cpu_rays = max(32, int(args.cpu_rays * 60 / args.steps) // 32 * 32)
    run_reference(args)
    return
  world = int(os.environ.get('WORLD_SIZE', '1'))
  if args.gpus > 1 and world == 1:
    # convenience: re-launch under torchrun when called directly with --gpus N
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
           '--master-addr', '127.0.0.1', '--master-port', '29517', os.path.abspath(__file__)] + sys.argv[1:]
    sys.exit(subprocess.call(cmd))
  run_ours(args)


if __name__ == '__main__':
  main()
