import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
for f in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us'):
  if os.path.exists(f):
    print(f, open(f).read().strip())
import bench
real = os.cpu_count
for th in (8, 16, 32, 64):
  os.cpu_count = lambda th=th: th
  t = time.time()
  r = bench.cpu_reference_step_rate('resnet50', 2, 1, 0)
  print(th, 'threads: %.3f transitions/s, %.1f s/step' % (r[0], r[2]), flush=True)
os.cpu_count = real
