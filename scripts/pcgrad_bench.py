"""Times t2r_pcgrad_project on the ResNet-50 critic's variable table (run on a B200):
  python scripts/pcgrad_bench.py > gpurun_out/pcgrad_bench.txt
Algorithmic bytes = (2 T reads + 1 write) x 4 B x parameters; compared with the measured HBM copy peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tensor2robot_b200 import nn  # noqa: E402
from tensor2robot_b200.research.qtopt import pcgrad  # noqa: E402
from tensor2robot_b200.research.qtopt import resnet_critic  # noqa: E402


def main():
  peaks = {}
  path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    peaks = json.load(open(path))
  from tensor2robot_b200 import engine
  from tensor2robot_b200.models import optimizers
  step = engine.CriticTrainStep(resnet_critic.ResNet50QCritic(), optimizers.MomentumOptimizer(1e-4, 0.9),
                                device=torch.device('cuda', 0), seed=0, world_size=1, rank=0)
  step.build(torch.zeros((2, 512, 640, 3), dtype=torch.uint8, device='cuda'), torch.zeros((2, 10), device='cuda'))
  vs = step.vs
  n = vs.flat.numel()
  opt = pcgrad.PCGrad(None)
  print('variables %d, parameters %.1f M' % (len(vs.trainable_variables()), n / 1e6))
  for tasks in (2, 4, 8):
    grads = torch.randn((tasks, n), dtype=torch.float32, device='cuda')
    for _ in range(3):
      opt.project(vs, grads)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    reps = 20
    for _ in range(reps):
      opt.project(vs, grads)
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / reps
    gbytes = (2 * tasks + 1) * 4 * n / 1e9
    print('tasks %d: %.3f ms, %.0f GB/s algorithmic (HBM copy peak %s GB/s)' % (
        tasks, ms, gbytes / (ms * 1e-3), peaks.get('hbm_copy_gbs', peaks.get('hbm_gbs', 'n/a'))))


if __name__ == '__main__':
  main()
