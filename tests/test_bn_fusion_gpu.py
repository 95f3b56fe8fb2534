"""The batch-norm fusions of the training step (nn._BnReluConvFn: masked / reduced data gradients,
operand-fused 1x1 convolutions) against the separate-node path they replace, on the ResNet-50 and
ResNet-18 critics: same weights, same inputs, all gradients.

The fused kernels use the same fmaf / bf16 rounding as t2r_bn_apply + the plain convolutions, so the forward
is bit-identical; gradients differ only by the summation order of fp32 atomics and of the
batch-norm reductions (layers/film_resnet_model.py:50-57, :283-340)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _record(name, **vals):
  path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'parity_r02.jsonl')
  try:
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, 'a') as f:
      f.write(json.dumps(dict(test=name, **vals)) + '\n')
  except OSError:
    pass


def _step(resnet_size, node, operand, b=4, size=96):
  from tensor2robot_b200 import nn
  from tensor2robot_b200.research.qtopt import resnet_critic
  old = nn.FUSE_BN_NODE, nn.FUSE_BN_OPERAND
  nn.FUSE_BN_NODE, nn.FUSE_BN_OPERAND = node, operand
  try:
    rng = np.random.RandomState(3)
    img = torch.from_numpy(rng.uniform(0, 1, (b, size, size, 3)).astype(np.float32)).cuda().to(torch.bfloat16)
    grasp = torch.from_numpy(rng.uniform(-1, 1, (b, 10)).astype(np.float32)).cuda()
    reward = torch.from_numpy((rng.uniform(size=(b, 1)) < 0.4).astype(np.float32)).cuda()
    vs = nn.VariableStore('cuda', seed=11)
    net = resnet_critic.ResNet50QCritic(resnet_size=resnet_size)
    with torch.no_grad(), nn.variable_store(vs):
      net.model((None, img[:2]), grasp[:2], is_training=False)
    vs.finalize()
    launches0 = __import__('tensor2robot_b200')._lib.launch_count()
    with nn.variable_store(vs):
      logits, _ = net.model((None, img), grasp, is_training=True)
      loss, q = nn.sigmoid_log_loss(logits, reward)
      vs.zero_grad()
      loss.backward()
    torch.cuda.synchronize()
    launches = __import__('tensor2robot_b200')._lib.launch_count() - launches0
    return (logits.float().cpu().numpy(), vs.flat_grad.clone().cpu().numpy(), vs.state_flat.clone().cpu().numpy(),
            launches, vs)
  finally:
    nn.FUSE_BN_NODE, nn.FUSE_BN_OPERAND = old


@pytest.mark.parametrize('resnet_size', [50, 18])
@pytest.mark.parametrize('operand', ['0', 'auto', 'all'])
def test_fused_bn_nodes_match_separate_nodes(resnet_size, operand):
  lo_ref, g_ref, st_ref, n_ref, vs = _step(resnet_size, False, '0')
  lo, g, st, n, _ = _step(resnet_size, True, operand)
  # forward: the same arithmetic, bit for bit (logits and the moving statistics)
  assert np.array_equal(lo, lo_ref), (lo[:4], lo_ref[:4])
  np.testing.assert_allclose(st, st_ref, rtol=1e-6, atol=1e-7)
  # gradients: per variable, relative l2 (fp32 atomics / reduction order only)
  worst, worst_name = 0.0, ''
  for name, v in vs.vars.items():
    if not v.trainable:
      continue
    a, r = g[v.offset:v.offset + v.numel], g_ref[v.offset:v.offset + v.numel]
    err = float(np.linalg.norm(a - r) / max(np.linalg.norm(r), 1e-20))
    if err > worst:
      worst, worst_name = err, name
  print('resnet%d operand=%s: launches %d -> %d, worst gradient rel-l2 %.3e (%s)' % (resnet_size, operand, n_ref, n,
                                                                                   worst, worst_name))
  _record('bn_fusion_vs_separate', resnet=resnet_size, operand=operand, launches_separate=n_ref, launches_fused=n,
          worst_grad_rel_l2=worst, worst_var=worst_name)
  assert n < n_ref            # the reduction (and, operand-fused, the apply) launches are gone
  assert worst < 2e-3, (worst, worst_name)
