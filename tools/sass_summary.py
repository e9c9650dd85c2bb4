"""Instruction histogram per kernel from `cuobjdump -sass` of the built objects: the Blackwell proof
(UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG = TMA loads/stores, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops) plus the usual suspects.  Writes profiles/r02_sass_summary.txt.

  python tools/sass_summary.py            (after multinerf_b200/build.py; no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, 'build', 'obj')
KEYS = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'LDTM', 'STTM', 'UTCBAR', 'UTCATOMSWS', 'SYNCS',
        'HMMA', 'RED', 'ATOM', 'MUFU', 'SHFL', 'LDG', 'STG', 'LDS', 'STS', 'BAR']


def main():
  out_path = os.path.join(ROOT, 'profiles', sys.argv[1] if len(sys.argv) > 1 else 'r02_sass_summary.txt')
  lines = ['# cuobjdump -sass build/obj/*.o (sm_100a), instruction counts per kernel; regenerate with tools/sass_summary.py',
           '# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG/UTMASTG = cp.async.bulk.tensor load/store, '
           'LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier', '']
  for obj in sorted(os.listdir(OBJ)):
    if not obj.endswith('.o'):
      continue
    txt = subprocess.run(['cuobjdump', '-sass', os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    kern, counts, total = None, {}, {}
    for ln in txt.splitlines():
      m = re.match(r'\s*Function : (\S+)', ln)
      if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
      m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', ln)
      if m and kern:
        op = m.group(1)
        total[kern] += 1
        base = op.split('.')[0]
        counts[kern][base] += 1
        if op.startswith('UTCHMMA') and '.2CTA' in op:
          counts[kern]['UTCHMMA.2CTA'] += 1
    for k in counts:
      name = subprocess.run(['c++filt', k], capture_output=True, text=True).stdout.strip()
      name = re.sub(r'\(.*', '', name)[:110]
      c = counts[k]
      shown = '  '.join(f'{key}={c[key]}' for key in KEYS if c[key])
      lines.append(f'{obj:14s} {name}\n{"":14s} {total[k]} instr: {shown}')
    lines.append('')
  os.makedirs(os.path.dirname(out_path), exist_ok=True)
  with open(out_path, 'w') as f:
    f.write('\n'.join(lines) + '\n')
  print(out_path)


if __name__ == '__main__':
  main()
