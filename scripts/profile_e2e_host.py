"""Host-side profile of the end-to-end training loop (train_eval_model on host batches) for one bench configuration:
where the wall clock of a step goes on the launching thread when the device step is shorter than the loop's period.

  python scripts/profile_e2e_host.py --config c4 [--steps 30]
"""
import cProfile
import io
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
  argv = sys.argv[1:]
  sys.argv = [sys.argv[0]] + argv + ['--no-cpu-baseline', '--no-extras']
  args = bench.parse_args()
  rt = bench.Runtime(args)
  model = bench.make_t2r_model(args, rt) if args.config in ('c4', 'c5') else None
  if model is None:
    from tensor2robot_b200.research.qtopt import t2r_models
    cls = t2r_models.ResNet50QCriticModel if args.model == 'resnet50' else \
        t2r_models.Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom
    model = cls(device=rt.dev)
  # the producer thread is invisible to cProfile: accumulate wall time around its two stages
  import collections
  import time
  from tensor2robot_b200.utils import train_eval
  spent = collections.Counter()

  def timed(cls, name):
    fn = getattr(cls, name)

    def wrapper(*a, **kw):
      t0 = time.perf_counter()
      try:
        return fn(*a, **kw)
      finally:
        spent[name] += time.perf_counter() - t0
        spent[name + ' calls'] += 1
    setattr(cls, name, wrapper)

  timed(train_eval.DeviceStager, 'stage')
  timed(train_eval.DeviceStager, '_fill')
  prof = cProfile.Profile()
  out = prof.runcall(bench.e2e_train_eval, rt, model, args.batch, args.steps, max(2, args.warmup))
  print('e2e: %.1f units/s, %.2f ms/step' % (out[0], out[3]))
  print('producer thread: ' + ', '.join('%s %.3f' % kv for kv in sorted(spent.items())))
  for key in ('tottime', 'cumulative'):
    s = io.StringIO()
    pstats.Stats(prof, stream=s).sort_stats(key).print_stats(28)
    print(s.getvalue()[:6000])
  rt.close()


if __name__ == '__main__':
  main()
