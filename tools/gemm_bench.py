"""Micro-benchmark of mnrf_gemm on the shapes of the 360 config (run on a B200).

  python tools/gemm_bench.py [--rows 1048576]

Prints ms / TFLOP/s / GB/s per (mode, N, K, features) so epilogue features can be costed in
isolation (each timing: 20 launches after 3 warm-ups, inputs larger than L2).
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multinerf_b200 import lib as L      # noqa: E402
from multinerf_b200 import ops           # noqa: E402


def timeit(fn, iters=20):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--rows', type=int, default=1 << 20)
  args = ap.parse_args()
  dev = torch.device('cuda:0')
  torch.manual_seed(0)
  for (M, N, K) in [(args.rows, 256, 256), (args.rows, 256, 512), (args.rows // 2, 1024, 1024)]:
    x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w_nk = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    w_kn = w_nk.t().contiguous()
    bias = torch.randn(N, device=dev) * 0.1
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    bits = torch.empty(M, max(N // 32, 4), device=dev, dtype=torch.int32)
    dy = (torch.randn(M, N, device=dev) * 0.1).bfloat16()
    dx = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
    xbits = torch.randint(-2**31, 2**31 - 1, (M, max(K // 32, 4)), device=dev, dtype=torch.int32)
    dw = torch.zeros(K, N, device=dev)
    cs = torch.zeros(K, device=dev)
    flops = 2.0 * M * N * K

    def report(name, ms, nbytes):
      print(f'M={M} N={N} K={K} {name:28s} {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s  '
            f'{nbytes / ms / 1e6:7.1f} GB/s', flush=True)

    ms = timeit(lambda: ops.gemm(L.GEMM_FWD, x, w_nk, y, m=M, n=N, k=K, act=L.ACT_RELU, bias=bias,
                                 maskbits=bits))
    report('fwd bias+relu+bits', ms, 2.0 * M * (N + K))
    ms = timeit(lambda: ops.gemm(L.GEMM_DGRAD, dy, w_kn, dx, m=M, n=K, k=N))
    report('dgrad plain', ms, 2.0 * M * (N + K))
    ms = timeit(lambda: ops.gemm(L.GEMM_DGRAD, dy, w_kn, dx, m=M, n=K, k=N, maskbits=xbits))
    report('dgrad bits', ms, 2.0 * M * (N + K))
    ms = timeit(lambda: ops.gemm(L.GEMM_DGRAD, dy, w_kn, dx, m=M, n=K, k=N, maskbits=xbits, colsum=cs))
    report('dgrad bits+colsum', ms, 2.0 * M * (N + K))
    ms = timeit(lambda: ops.gemm(L.GEMM_WGRAD, x, dy, dw, m=K, n=N, k=M))
    report('wgrad', ms, 2.0 * M * (N + K))
    ms = timeit(lambda: ops.colsum(dy, N, cs[:N] if N <= K else torch.zeros(N, device=dev)))
    report('colsum kernel', ms, 2.0 * M * N)
    del x, y, dy, dx, bits, xbits


if __name__ == '__main__':
  main()
