#!/usr/bin/env python
"""Per-kernel cost of a small ray shard against the full batch: where a strong-scaled step loses efficiency.

usage: python tools/compare_launches.py <full launches.csv> <shard launches.csv> <shard factor> > profiles/<tag>.txt

Both inputs are `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch lists of
`bench.py --steps 1 --warmup 3 --no_graph` (the shard one with --batch_size 16384/<factor>); the last full train
step of each is compared kernel by kernel (same launch order).
"""
import collections
import csv
import re
import sys


def last_step(path):
  lines = [l for l in open(path) if not l.startswith('==')]
  rows = [(r['Kernel Name'], float(r['Metric Value'].replace(',', '')))
          for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
  idx = [i for i, (n, _) in enumerate(rows) if 'sample_level' in n][::3]
  return rows[idx[-2]:idx[-1]]


def main():
  full, shard, factor = last_step(sys.argv[1]), last_step(sys.argv[2]), float(sys.argv[3])
  assert len(full) == len(shard), (len(full), len(shard))
  agg = collections.OrderedDict()
  for (n, t), (n2, t2) in zip(full, shard):
    assert n[:30] == n2[:30], (n, n2)
    k = re.sub(r'\(.*', '', n).replace('void ', '')[:52]
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += t
    a[2] += t2
  tf, ts = sum(t for _, t in full), sum(t for _, t in shard)
  print(f'# one 360.gin train step, kernel by kernel: full batch (16384 rays) vs a 1/{factor:g} shard on ONE GPU')
  print(f'# serialised totals (ncu, burst clocks): full {tf / 1e6:.3f} ms, full/{factor:g} = {tf / factor / 1e6:.3f} ms, '
        f'shard {ts / 1e6:.3f} ms -> {100 * (ts / (tf / factor) - 1):.1f} % over the ideal')
  print('# launches   full_us   full/f_us   shard_us   excess_us  kernel')
  for k, v in sorted(agg.items(), key=lambda kv: -(kv[1][2] - kv[1][1] / factor)):
    print(f'{v[0]:6d} {v[1] / 1e3:11.1f} {v[1] / factor / 1e3:10.1f} {v[2] / 1e3:10.1f} {v[2] / 1e3 - v[1] / factor / 1e3:10.1f}  {k}')


if __name__ == '__main__':
  main()
