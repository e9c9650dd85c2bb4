"""Thread sweep of the CPU arm (oracle restatement of the 360.gin train step, fp32 torch-CPU) on the GPU
box's host: picks the thread count bench.py pins in CPU_THREADS_DEFAULT.  Each point runs in its own
process (the OpenMP pool size is fixed at start-up).  Output -> profiles/r02_cpu_sweep.txt

  python tools/cpu_sweep.py [rays=1024] [threads ...]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
  rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
  threads = [int(t) for t in sys.argv[2:]] or [16, 32, 64, 128]
  cores = os.cpu_count()
  rows = [f'# CPU arm thread sweep: oracle 360.gin train step, {rays} rays per step, 1 warm-up + 2 timed steps, '
          f'host with {cores} logical cores', '# threads  rays/s  s/step']
  for t in threads:
    if t > cores:
      continue
    env = dict(os.environ, MNRF_CPU_THREADS=str(t))
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2',
                        '--warmup', '1', '--cpu_rays', str(rays)], capture_output=True, text=True, env=env)
    try:
      j = json.loads(r.stdout.strip().splitlines()[-1])
      rows.append(f'{t:8d}  {j["value"]:8.1f}  {j["ms_per_step"] / 1e3:6.2f}')
    except Exception:  # pylint: disable=broad-except
      rows.append(f'{t:8d}  failed: {r.stderr[-200:]}')
    print(rows[-1], flush=True)
  out = os.path.join(ROOT, 'gpurun_out', 'cpu_sweep.txt')
  os.makedirs(os.path.dirname(out), exist_ok=True)
  with open(out, 'w') as f:
    f.write('\n'.join(rows) + '\n')


if __name__ == '__main__':
  main()
