"""Per-kernel device-time shares of ONE training step from an `ncu --metrics gpu__time_duration.sum` launch list.

  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-e2e --no-extras --no-cpu-baseline
  python scripts/launch_shares.py gpurun_out/launches.csv [steps_in_capture] > profiles/r02_launch_shares_bench_default.txt

The capture contains the build pass, the warm-up step and the timed steps; the LAST `steps_in_capture`-th of the launches
between the first and last optimizer kernel are attributed to one step by dividing by the number of optimizer launches."""
import collections
import csv
import re
import sys


def main():
  path = sys.argv[1]
  rows = []
  with open(path) as f:
    for line in f:
      if line.startswith('"ID"'):
        break
    for r in csv.reader(f):
      if len(r) >= 15 and r[12] == 'gpu__time_duration.sum':
        rows.append((r[4], float(r[14])))
  opt = [i for i, (n, _) in enumerate(rows) if 'optimizer_kernel' in n or 'momentum_kernel' in n]
  if len(opt) < 2:
    raise SystemExit('need at least two optimizer launches in the capture')
  # one step = launches after the previous optimizer kernel up to and including this one; use the last step
  lo, hi = opt[-2] + 1, opt[-1] + 1
  step = rows[lo:hi]
  agg = collections.OrderedDict()
  for name, ns in step:
    short = re.sub(r'\(.*$', '', name).replace('void ', '').replace('t2r::', '')
    d = agg.setdefault(short, [0, 0.0])
    d[0] += 1
    d[1] += ns
  total = sum(v[1] for v in agg.values())
  for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-72s n=%4d %9.3f ms %5.1f%%' % (name[:72], n, ns / 1e6, 100 * ns / total))
  print('total %.3f ms over %d launches (cold-cache, serialised, one step)' % (total / 1e6, len(step)))


if __name__ == '__main__':
  main()
