import sys, time, torch, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tensor2robot_b200.utils import tensorspec_utils, train_eval
sys.argv = ['x', '--config', 'c4', '--no-cpu-baseline', '--no-extras']
args = bench.parse_args()
rt = bench.Runtime(args)
model = bench.make_t2r_model(args, rt)
pre = model.preprocessor
sets = []
for _ in range(2):
  f = tensorspec_utils.make_random_numpy(pre.get_in_feature_specification('train'), args.batch)
  l = tensorspec_utils.make_random_numpy(pre.get_in_label_specification('train'), args.batch)
  merged = tensorspec_utils.TensorSpecStruct([('f/' + k, v) for k, v in tensorspec_utils.flatten_spec_structure(f).items()])
  for k, v in tensorspec_utils.flatten_spec_structure(l).items():
    merged['l/' + k] = v
  sets.append(merged)
for k, v in sets[0].items():
  print(k, v.dtype, v.shape, v.nbytes)
import collections
spent = collections.Counter()
def timed(obj, name, label):
  fn = getattr(obj, name)
  def wrapper(*a, **kw):
    t = time.perf_counter()
    try:
      return fn(*a, **kw)
    finally:
      spent[label] += time.perf_counter() - t
  setattr(obj, name, wrapper)
timed(train_eval.DeviceStager, '_fill', 'fill')
timed(torch.cuda.Event, 'synchronize', 'slot event sync')
_to = torch.Tensor.to
def to(self, *a, **kw):
  t = time.perf_counter()
  try:
    return _to(self, *a, **kw)
  finally:
    spent['to(device)'] += time.perf_counter() - t
torch.Tensor.to = to
stager = train_eval.DeviceStager(rt.dev, depth=4)
cs = torch.cuda.current_stream(rt.dev)
for i in range(8):
  stager.stage(sets[i % 2], cs)
torch.cuda.synchronize()
spent.clear()
t0 = time.perf_counter()
for i in range(40):
  stager.stage(sets[i % 2], cs)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print({k: round(v / 40 * 1e3, 2) for k, v in spent.items()}, 'ms per call')
print('stage alone: %.2f ms per call (host), %.2f ms per call incl. device completion' % ((t1 - t0) / 40 * 1e3, (t2 - t0) / 40 * 1e3))
