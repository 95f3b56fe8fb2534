"""The batch-norm fusions of the training step (nn._BnReluConvFn: one autograd node for BN + ReLU + the consuming
convolutions, operand-fused 1x1 kernels, masked / reduced data gradients) against the separate-node path they replace.

Block level (one ResNet v2 bottleneck / building block, layers/film_resnet_model.py:166-223, :283-340): same weights
and inputs, outputs / input gradients / every parameter gradient / moving statistics must agree up to the summation
order of fp32 atomics (a batch-norm statistic that differs in its last bit can flip the bf16 rounding of a few
activations).  Network level (ResNet-50 / ResNet-18 critics): a randomly initialised 50-layer training-mode BN
network amplifies such flips chaotically, so there the check is that the fused graph launches
fewer kernels and produces the same loss to within that conditioning."""
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


def _run_block(kind, node, operand, projection, strides, filters, hw, b=4):
  from tensor2robot_b200 import _lib, nn
  from tensor2robot_b200.layers import film_resnet_model as rm
  old = nn.FUSE_BN_NODE, nn.FUSE_BN_OPERAND
  nn.FUSE_BN_NODE, nn.FUSE_BN_OPERAND = node, operand
  try:
    bottleneck = kind == 'bottleneck'
    block = rm._bottleneck_block_v2 if bottleneck else rm._building_block_v2
    c_out = 4 * filters if bottleneck else filters
    c_in = c_out // 2 if projection else c_out
    rng = np.random.RandomState(7)
    x0 = torch.from_numpy(rng.standard_normal((b, hw, hw, c_in)).astype(np.float32)).cuda().to(torch.bfloat16)
    vs = nn.VariableStore('cuda', seed=5)

    def call(x, training):
      namer = rm._Namer()
      proj = None
      if projection:
        proj = lambda t: rm.conv2d_fixed_padding(t, c_out, 1, strides, namer, 1e-4)
        proj.fused_args = (c_out, strides)
      return block(x, filters, training, proj, strides, namer, 1e-4)

    with torch.no_grad(), nn.variable_store(vs):
      call(x0, False)
    vs.finalize()
    x = x0.clone().requires_grad_(True)
    launches0 = _lib.launch_count()
    with nn.variable_store(vs):
      y = call(x, True)
      wts = torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32)).cuda()
      loss = (y.float() * wts).sum()
      vs.zero_grad()
      loss.backward()
    torch.cuda.synchronize()
    return dict(y=y.detach().float().cpu().numpy(), dx=x.grad.float().cpu().numpy(), grad=vs.flat_grad.cpu().numpy(),
                state=vs.state_flat.cpu().numpy(), launches=_lib.launch_count() - launches0, vs=vs)
  finally:
    nn.FUSE_BN_NODE, nn.FUSE_BN_OPERAND = old


def _rel_l2(a, b):
  return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))


@pytest.mark.parametrize('kind,projection,strides,filters,hw', [
    ('bottleneck', False, 1, 64, 30),    # identity shortcut: BN routes the residual gradient (passthrough)
    ('bottleneck', True, 1, 64, 30),     # first block of layer 1: projection + first conv share the BN
    ('bottleneck', True, 2, 128, 30),    # strided projection, 3x3 stride 2
    ('bottleneck', False, 1, 256, 15),   # 1024 -> 256: not operand-fused in 'auto'
    ('building', False, 1, 64, 24),      # ResNet-18 block: 3x3 consumers (halo kernels), node fusion only
    ('building', True, 2, 128, 24),
])
@pytest.mark.parametrize('operand', ['0', 'auto', 'all'])
def test_fused_block_matches_separate_nodes(kind, projection, strides, filters, hw, operand):
  ref = _run_block(kind, False, '0', projection, strides, filters, hw)
  got = _run_block(kind, True, operand, projection, strides, filters, hw)
  scale = np.abs(ref['y']).max()
  bad = np.abs(got['y'] - ref['y']) > 2 * 2.0**-8 * scale          # more than 2 bf16 ulp of the output scale
  errs = dict(y_bad_frac=float(bad.mean()), y=_rel_l2(got['y'], ref['y']), dx=_rel_l2(got['dx'], ref['dx']),
              grad=_rel_l2(got['grad'], ref['grad']), state=_rel_l2(got['state'], ref['state']))
  print('%s proj=%s s=%d f=%d operand=%s: launches %d -> %d, %s' % (kind, projection, strides, filters, operand,
                                                                    ref['launches'], got['launches'], errs))
  _record('bn_fusion_block', kind=kind, projection=projection, strides=strides, filters=filters, operand=operand,
          launches_separate=ref['launches'], launches_fused=got['launches'], **errs)
  assert errs['y_bad_frac'] < 1e-3 and errs['y'] < 2e-3
  assert errs['state'] < 1e-5
  assert errs['dx'] < 5e-3 and errs['grad'] < 5e-3
  assert got['launches'] <= ref['launches']


def _critic_step(resnet_size, node, operand, b=4, size=96):
  from tensor2robot_b200 import _lib, nn
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
    launches0 = _lib.launch_count()
    with nn.variable_store(vs):
      logits, _ = net.model((None, img), grasp, is_training=True)
      loss, _ = nn.sigmoid_log_loss(logits, reward)
      vs.zero_grad()
      loss.backward()
    torch.cuda.synchronize()
    return float(loss.detach()), vs.flat_grad.cpu().numpy(), _lib.launch_count() - launches0
  finally:
    nn.FUSE_BN_NODE, nn.FUSE_BN_OPERAND = old


@pytest.mark.parametrize('resnet_size', [50, 18])
def test_fused_critic_step_runs(resnet_size):
  loss_ref, g_ref, n_ref = _critic_step(resnet_size, False, '0')
  loss, g, n = _critic_step(resnet_size, True, 'auto')
  print('resnet%d: launches %d -> %d, loss %.5f vs %.5f' % (resnet_size, n_ref, n, loss, loss_ref))
  assert np.isfinite(g).all() and np.abs(g).max() > 0
  assert abs(loss - loss_ref) < 5e-2 * max(1.0, abs(loss_ref))
  assert n <= n_ref          # ResNet-18 has no 1x1 reductions to operand-fuse: same launches, one node fewer per block
