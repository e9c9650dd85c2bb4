#!/usr/bin/env python
"""Summarise gpurun_out/{launches.csv, <name>.ncu-rep} into profiles/<tag>_*.txt (committed).

usage: python tools/summarize_profile.py <tag> [ncu-rep]
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = sys.argv[2] if len(sys.argv) > 2 else None
out_dir = os.path.join(ROOT, 'profiles')
os.makedirs(out_dir, exist_ok=True)

lp = os.path.join(ROOT, 'gpurun_out', 'launches.csv')
if os.path.exists(lp):
  lines = [l for l in open(lp) if not l.startswith('==')]
  rows = [(r['Kernel Name'], float(r['Metric Value'].replace(',', '')))
          for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
  names = [n for n, _ in rows]
  idx = [i for i, n in enumerate(names) if 'sample_level' in n]
  nlev = 3
  starts = idx[::nlev]
  s0, s1 = starts[-2], starts[-1]
  step = rows[s0:s1]
  agg = collections.OrderedDict()
  for n, t in step:
    k = re.sub(r'\(.*', '', n)[:80]
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += t
  tot = sum(v[1] for v in agg.values())
  with open(os.path.join(out_dir, f'{tag}_launches_summary.txt'), 'w') as f:
    f.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 1 --warmup 3\n')
    f.write(f'# one full train step (360.gin, 16384 rays), {len(step)} launches, serialised total '
            f'{tot / 1e6:.3f} ms (cold-cache, serialised: compare SHARES)\n')
    f.write('#   time_ms  share  launches  kernel\n')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
      f.write(f'{v[1] / 1e6:10.3f} {100 * v[1] / tot:6.1f}% {v[0]:5d}  {k}\n')
  with open(os.path.join(out_dir, f'{tag}_launches_step.csv'), 'w') as f:
    f.write('order,kernel,duration_ns\n')
    for i, (n, t) in enumerate(step):
      f.write(f'{i},"{re.sub(chr(34), "", n)[:120]}",{t:.0f}\n')
  print(open(os.path.join(out_dir, f'{tag}_launches_summary.txt')).read())

if rep:
  raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  r = list(csv.reader(raw.splitlines()))
  hdr, units = r[0], r[1]
  want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
          'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
          'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max',
          'smsp__cycles_active.avg', 'launch__shared_mem_per_block_dynamic']
  idx = [i for i, h in enumerate(hdr) if h in want]
  with open(os.path.join(out_dir, f'{tag}_{os.path.basename(rep).replace(".ncu-rep", "")}_ncu.txt'), 'w') as f:
    f.write(f'# ncu --set full --clock-control none --import-source on; source: {os.path.basename(rep)}\n')
    for row in r[2:]:
      f.write('----\n')
      for i in idx:
        f.write(f'{hdr[i]} [{units[i]}] = {row[i][:100]}\n')
  print('wrote ncu summary')

# per-launch DRAM traffic of the dominant kernel over one whole step (light metric set, every GEMM
# launch) -> profiles/<tag>_gemm_traffic.txt and bench.py's roofline.traffic
tp = os.path.join(ROOT, 'gpurun_out', 'gemm_traffic.csv')
if os.path.exists(tp):
  import json
  lines = [l for l in open(tp) if l.startswith('"')]
  per = collections.OrderedDict()
  for r in csv.DictReader(lines):
    d = per.setdefault(int(r['ID']), {'kernel': re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '')})
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'nsecond': 1.0, 'usecond': 1e3, 'msecond': 1e6,
             'second': 1e9}.get(unit, 1.0)
    d[r['Metric Name']] = v * scale
  rows = list(per.values())
  tot = [r['dram__bytes_read.sum'] + r['dram__bytes_write.sum'] for r in rows]
  with open(os.path.join(out_dir, f'{tag}_gemm_traffic.txt'), 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,'
            'sm__pipe_tensor_cycles_active...,lts__t_bytes.sum --clock-control none\n')
    f.write(f'# every tensor-core launch (gemm_tc_kernel + mlp_chain_kernel) of one 360.gin train step (16384 rays): {len(rows)} launches, '
            f'DRAM {sum(tot) / 1e9:.2f} GB per step, {sum(tot) / len(tot) / 1e9:.4f} GB per launch on average\n')
    f.write('#  i  kernel<mode,ctas,stages,out>            time_us  dram_rd_MB  dram_wr_MB  L2_GB  tensor_pipe_%\n')
    for i, r in enumerate(rows):
      f.write(f"{i:4d}  {r['kernel'][:38]:38s} {r['gpu__time_duration.sum'] / 1e3:9.1f} "
              f"{r['dram__bytes_read.sum'] / 1e6:10.1f} {r['dram__bytes_write.sum'] / 1e6:10.1f} "
              f"{r.get('lts__t_bytes.sum', 0) / 1e9:7.2f} "
              f"{r.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):8.1f}\n")
  with open(os.path.join(out_dir, 'gemm_tc_traffic.json'), 'w') as f:
    json.dump({'kernel': 'gemm_tc_kernel + mlp_chain_kernel', 'launches': len(tot), 'dram_bytes_per_launch': sum(tot) / len(tot),
               'dram_bytes_per_step': sum(tot),
               'source': f'profiles/{tag}_gemm_traffic.txt (ncu dram__bytes_read.sum + dram__bytes_write.sum, '
                         'every tensor-core launch of one 360.gin train step)'}, f, indent=1)
  print('wrote gemm_tc_traffic.json:', sum(tot) / len(tot) / 1e9, 'GB per launch,', sum(tot) / 1e9, 'GB per step')
