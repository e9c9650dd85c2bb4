"""Event timeline of the chained-trunk kernel (knob build of the library: tools/build_variant notes in
DESIGN.md).  CTA 0 (leader of pair 0) logs clock64 stamps for its MMA thread, its hand-off warp and one
epilogue thread; this prints the timeline of the first units in SM cycles relative to the first event.

  MNRF_LIB=$PWD/multinerf_b200/libmnrf_b200_knobs.so python tools/chain_trace.py [--train]
"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinerf_b200 import lib as L, ops  # noqa: E402

NAMES = {10: 'mma  wait act_ready', 11: 'mma  act_ready seen', 12: 'mma  phase issued',
         20: 'epi  wait acc_full', 21: 'epi  acc_full seen', 22: 'epi  buf_free seen', 23: 'epi  compute done',
         24: 'epi  fenced', 25: 'epi  after bar', 30: 'w3   after bar', 31: 'w3   arrived act_ready',
         32: 'w3   buf_free arrived'}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--train', action='store_true')
  ap.add_argument('--m', type=int, default=16384 * 64)
  a = ap.parse_args()
  L.require_device()
  M, D, F, W = a.m, 4, 512, 256
  dev = 'cuda'
  g = torch.Generator(device=dev).manual_seed(0)
  bf = torch.bfloat16
  feat = torch.randn(M, F, device=dev, generator=g).to(bf)
  ws = [(torch.randn(W, F if i == 0 else W, device=dev, generator=g) * math.sqrt(2.0 / (F if i == 0 else W))).to(bf)
        for i in range(D)]
  bs = [torch.zeros(W, device=dev) for _ in range(D)]
  acts = [torch.empty(M, W, device=dev, dtype=bf) for _ in range(D)]
  bits = [torch.empty(M, W // 32, device=dev, dtype=torch.int32) for _ in range(D)]
  head = torch.empty(M, device=dev)
  hw = torch.randn(W, device=dev, generator=g)
  hb = torch.zeros(1, device=dev)
  layers = []
  for i in range(D):
    ly = dict(w=ws[i], bias=bs[i])
    ly.update(dict(n_stream=F // 64) if i == 0 else dict(n_res=4))
    if a.train or i == D - 1:
      ly['out'] = acts[i]
    if a.train:
      ly['maskbits'] = bits[i]
    layers.append(ly)
  desc = ops.chain_desc(L.CHAIN_FWD, M, layers, stream=feat, stream_cols=F, head_w=hw, head_b=hb, head_out=head)
  for _ in range(3):
    ops.mlp_chain(desc)
  torch.cuda.synchronize()
  trace = torch.zeros(1 + 3 * 4000, dtype=torch.int64, device=dev)
  os.environ['MNRF_CHAIN_TRACE'] = str(trace.data_ptr())
  ops.mlp_chain(desc)
  torch.cuda.synchronize()
  os.environ.pop('MNRF_CHAIN_TRACE')
  t = trace.cpu()
  n = int(t[0])
  ev = sorted((int(t[3 + 3 * i]), int(t[1 + 3 * i]), int(t[2 + 3 * i])) for i in range(min(n, 4000)))
  if not ev:
    print('no events: is MNRF_LIB the knob build?')
    return
  t0 = ev[0][0]
  print(f'# {n} events; columns: cycles since first event, event, unit/layer/block')
  for clk, tag, key in ev:
    unit, rest = divmod(key, 100)
    j, X = divmod(rest, 10)
    print(f'{clk - t0:9d}  {NAMES.get(tag, tag):24s} unit {unit:4d} layer {j} block {"AB"[X]}')


if __name__ == '__main__':
  main()
