"""Layer-by-layer comparison of the engine's Grasping44 with the oracle (debug aid, GPU only).

usage: python tests/parity_trace.py [train|eval]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import qtopt_networks as oracle, tf_ops
from tensor2robot_b200 import nn
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_qtopt_networks_gpu import _inputs, _build_engine, _rel_l2

training = (sys.argv[1] if len(sys.argv) > 1 else 'eval') == 'train'
b = 4
img, grasp, reward = _inputs(b)
variables = oracle.init_variables(seed=3)
for k in variables:
  if k.endswith('/weights'):
    variables[k] = (variables[k] * 8).astype(np.float32)
img_t = torch.from_numpy(img).cuda().to(torch.bfloat16)
grasp_t = torch.from_numpy(grasp).cuda()
vs, net = _build_engine(img_t, grasp_t, variables)
nn.TRACE = []
with torch.no_grad(), nn.variable_store(vs):
  net.model((None, img_t), grasp_t, is_training=training)
eng = [(op, name, y.float().cpu().numpy()) for op, name, y in nn.TRACE]
tf_ops.TRACE = []
with torch.no_grad():
  oracle.model(oracle.to_torch(variables, False), img_t.float().cpu(), torch.from_numpy(grasp), training)
orc = [(op, y.numpy()) for op, y in tf_ops.TRACE]
# the oracle sums 7 dense blocks where the engine runs one fc32: merge them
merged, i = [], 0
while i < len(orc):
  op, y = orc[i]
  if op == 'dense' and y.shape[-1] == 256:
    acc = y
    j = i + 1
    while j < len(orc) and orc[j][0] == 'dense' and orc[j][1].shape[-1] == 256:
      acc = acc + orc[j][1]; j += 1
    merged.append(('dense', acc)); i = j
  else:
    merged.append((op, y)); i += 1
# engine order: image tower ..., fc32, bn, conv(fcgrasp2) ... ; oracle has the same order
print(len(eng), len(merged))
eng = [e for e in eng if e[0] != 'add_context']
for (eop, name, ey), (oop, oy) in zip(eng, merged):
  if eop == 'bn':
    oy = np.maximum(oy, 0)
  ey = ey.reshape(oy.shape) if ey.size == oy.size else ey
  print('%-12s %-12s %-70s %-22s rel_l2 %.3e  |ref| %.3e' % (eop, oop, name[-70:], str(ey.shape), _rel_l2(ey, oy) if ey.shape == oy.shape else -1, np.abs(oy).mean()))
