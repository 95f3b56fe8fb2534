"""cProfile of the launching thread over device-resident training steps of one bench configuration (no input pipeline):
which Python frames the host time of a step goes to when the step is launch-bound (BC-Z: ~300 launches per 13 ms).

  python scripts/profile_step_host.py --config c4 [--steps 30]
"""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from tensor2robot_b200.utils import tensorspec_utils  # noqa: E402


def main():
  sys.argv = [sys.argv[0]] + sys.argv[1:] + ['--no-cpu-baseline', '--no-extras', '--no-e2e']
  args = bench.parse_args()
  rt = bench.Runtime(args)
  model = bench.make_t2r_model(args, rt)
  pre = model.preprocessor

  def to_dev(spec):
    flat = tensorspec_utils.flatten_spec_structure(tensorspec_utils.make_random_numpy(spec, args.batch))
    return tensorspec_utils.TensorSpecStruct([(k, torch.from_numpy(np.ascontiguousarray(v)).to(rt.dev)) for k, v in flat.items()])

  f, l = to_dev(pre.get_in_feature_specification('train')), to_dev(pre.get_in_label_specification('train'))
  clone = lambda st: tensorspec_utils.TensorSpecStruct(list(tensorspec_utils.flatten_spec_structure(st).items()))

  def one():
    features, labels = pre.preprocess(clone(f), clone(l) if len(l) else None, 'train')
    return model.train_step(features, labels)

  for _ in range(5):
    one()
  torch.cuda.synchronize()
  prof = cProfile.Profile()
  t0 = time.perf_counter()
  prof.enable()
  for _ in range(args.steps):
    one()
  prof.disable()
  t_host = time.perf_counter() - t0
  torch.cuda.synchronize()
  t_all = time.perf_counter() - t0
  print('%d steps: host loop %.2f ms/step, with the device drained %.2f ms/step' % (
      args.steps, t_host / args.steps * 1e3, t_all / args.steps * 1e3))
  for key in ('tottime', 'cumulative'):
    s = io.StringIO()
    pstats.Stats(prof, stream=s).sort_stats(key).print_stats(32)
    print(s.getvalue()[:7000])
  rt.close()


if __name__ == '__main__':
  main()
