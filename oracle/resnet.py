"""ORACLE (test infrastructure, never imported by the product): the reference's ResNet v2 (+FiLM)
and our ResNet-50 Q-critic composition restated on torch CPU fp32.

Follows layers/film_resnet_model.py:39-57 (batch_norm), :60-105 (conv2d_fixed_padding),
:108-115 (_apply_film), :166-223 / :283-340 (v2 blocks), :343-388 (block_layer), :525-629
(Model.__call__) and layers/resnet.py:31-62,98-143.  Variables carry the names tf.layers would
auto-generate under 'resnet_model/' (conv2d, conv2d_1, ..., batch_normalization, ...), kernels HWIO.
PARITY UNPINNED: the reference only tests endpoint names/shapes for this network.
"""
import collections

import numpy as np
import torch

from oracle import tf_ops

BN_DECAY, BN_EPS = 0.997, 1e-5
BLOCK_SIZES = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3], 50: [3, 4, 6, 3], 101: [3, 4, 23, 3]}


class Namer(object):

  def __init__(self):
    self.counts = {}

  def __call__(self, base):
    n = self.counts.get(base, 0)
    self.counts[base] = n + 1
    return base if n == 0 else '%s_%d' % (base, n)


class _Builder(object):
  """Walks the graph once; in 'init' mode it creates variables, otherwise it computes."""

  def __init__(self, variables, training, scope, updates=None, rng=None):
    self.v, self.training, self.scope, self.updates, self.rng = variables, training, scope, updates, rng
    self.namer = Namer()

  def conv(self, x, filters, k, strides, residual=None):
    name = self.scope + self.namer('conv2d') + '/kernel'
    if self.rng is not None and name not in self.v:
      cin = x.shape[-1]
      std = tf_ops.variance_scaling_std(k * k * cin)
      self.v[name] = torch.tensor(np.clip(self.rng.normal(0, std, (k, k, cin, filters)), -2 * std, 2 * std),
                                  dtype=torch.float32)
    return tf_ops.conv2d_fixed_padding(x, self.v[name], strides, residual)

  def bn(self, x, relu=False, film=None):
    name = self.scope + self.namer('batch_normalization')
    c = x.shape[-1]
    if self.rng is not None and name + '/beta' not in self.v:
      self.v[name + '/gamma'] = torch.ones(c)
      self.v[name + '/beta'] = torch.zeros(c)
      self.v[name + '/moving_mean'] = torch.zeros(c)
      self.v[name + '/moving_variance'] = torch.ones(c)
    y = tf_ops.batch_norm(x, self.v, name, self.training, BN_DECAY, BN_EPS, True, self.updates)
    if film is not None:                                  # _apply_film, film_resnet_model.py:108-115
      gamma, beta = film[:, None, None, :c], film[:, None, None, c:]
      y = (1 + gamma) * y + beta
    return tf_ops.relu(y) if relu else y


def _block_v2(b, x, filters, bottleneck, project, strides, film):
  shortcut = x
  x = b.bn(x, relu=True)
  if project:
    shortcut = b.conv(x, filters * 4 if bottleneck else filters, 1, strides)
  if bottleneck:
    x = b.conv(x, filters, 1, 1)
    x = b.bn(x, relu=True)
    x = b.conv(x, filters, 3, strides)
    x = b.bn(x, relu=True, film=film)
    x = b.conv(x, 4 * filters, 1, 1, residual=shortcut)
  else:
    x = b.conv(x, filters, 3, strides)
    x = b.bn(x, relu=True, film=film)
    x = b.conv(x, filters, 3, 1, residual=shortcut)
  return x


def _block_v1(b, x, filters, bottleneck, project, strides, film):
  """film_resnet_model.py:121-168 / :220-276: post-activation blocks; the projection shortcut has its own batch norm,
  FiLM modulates the last batch norm's output, the ReLU follows the shortcut add."""
  shortcut = x
  if project:
    shortcut = b.bn(b.conv(x, filters * 4 if bottleneck else filters, 1, strides))
  if bottleneck:
    x = b.bn(b.conv(x, filters, 1, 1), relu=True)
    x = b.bn(b.conv(x, filters, 3, strides), relu=True)
    x = b.bn(b.conv(x, 4 * filters, 1, 1), film=film)
  else:
    x = b.bn(b.conv(x, filters, 3, strides), relu=True)
    x = b.bn(b.conv(x, filters, 3, 1), film=film)
  return tf_ops.relu(tf_ops._store(x + shortcut))


def block_layers(b, x, resnet_size, first, last, films=None, num_filters=64, end_points=None, version=2):
  sizes = BLOCK_SIZES[resnet_size]
  bottleneck = resnet_size >= 50
  strides = [1, 2, 2, 2]
  block = _block_v2 if version == 2 else _block_v1
  for i in range(first, last):
    f = num_filters * 2**i
    for j in range(sizes[i]):
      film = films[i][j] if films is not None and films[i] is not None else None
      x = block(b, x, f, bottleneck, j == 0, strides[i] if j == 0 else 1, film)
    if end_points is not None:
      end_points['block_layer%d' % (i + 1)] = x
  return x


def stem(b, x, num_filters=64, kernel_size=7, end_points=None):
  x = b.conv(x, num_filters, kernel_size, 2)
  if end_points is not None:
    end_points['initial_conv'] = x
  x = tf_ops.max_pool(x, 3, 2, 'SAME')
  if end_points is not None:
    end_points['initial_max_pool'] = x
  return x


def resnet_model(variables, images, training, num_classes, resnet_size=50, films=None, updates=None,
                 rng=None, scope='resnet_model/', end_points=None, version=2):
  """layers/resnet.py:147-209 + Model.__call__.  rng != None creates missing variables.  version 1: BN + ReLU after
  the stem convolution, post-activation blocks, no final BN (film_resnet_model.py:565-571, 603-610)."""
  b = _Builder(variables, training, scope, updates, rng)
  if version == 1:
    x = b.conv(images, 64, 7, 2)
    x = b.bn(x, relu=True)
    x = tf_ops.max_pool(x, 3, 2, 'SAME')
  else:
    x = stem(b, images, end_points=end_points)
  x = block_layers(b, x, resnet_size, 0, 4, films, end_points=end_points, version=version)
  if version == 2:
    x = b.bn(x, relu=True)
  x = tf_ops._store(x.mean((1, 2)))
  name = scope + b.namer('dense')
  if rng is not None and name + '/kernel' not in variables:
    k = x.shape[-1]
    lim = np.sqrt(6.0 / (k + num_classes))
    variables[name + '/kernel'] = torch.tensor(rng.uniform(-lim, lim, (k, num_classes)), dtype=torch.float32)
    variables[name + '/bias'] = torch.zeros(num_classes)
  return tf_ops.dense(x, variables[name + '/kernel'], variables[name + '/bias'])


# ---- our ResNet-50 Q-critic composition (tensor2robot_b200/research/qtopt/resnet_critic.py) ----
GRASP_PARAM_NAMES = collections.OrderedDict([
    ('fcgrasp_wv', (0, 3)), ('fcgrasp_vr', (3, 2)), ('fcgrasp_gripper_close', (5, 1)),
    ('fcgrasp_gripper_open', (6, 1)), ('fcgrasp_terminate_episode', (7, 1)),
    ('fcgrasp_gripper_closed', (8, 1)), ('fcgrasp_height_to_bottom', (9, 1))])
HEAD_BN_DECAY, HEAD_BN_EPS, L2 = 0.9997, 0.001, 0.00007


def critic(variables, image, grasp_params, training, resnet_size=50, merge_after=3, updates=None, rng=None,
           scope='ResNet50QCritic/', end_points=None):
  """Returns logits [M,1]; rng != None initialises missing variables (truncated_normal .01 heads)."""
  v, p = variables, scope
  tile = grasp_params.dim() == 3
  a = grasp_params.shape[1] if tile else 1
  if tile:
    grasp_params = grasp_params.reshape(-1, grasp_params.shape[2])

  def ensure(name, shape, std=0.01, const=None):
    if rng is not None and name not in v:
      v[name] = (torch.full(shape, float(const)) if const is not None else
                 torch.tensor(np.clip(rng.normal(0, std, shape), -2 * std, 2 * std), dtype=torch.float32))

  def head_bn(x, name, scale=True):
    c = x.shape[-1]
    ensure(name + '/beta', (c,), const=0)
    if scale:
      ensure(name + '/gamma', (c,), const=1)
    ensure(name + '/moving_mean', (c,), const=0)
    ensure(name + '/moving_variance', (c,), const=1)
    return tf_ops.batch_norm(x, v, name, training, HEAD_BN_DECAY, HEAD_BN_EPS, scale, updates)

  b = _Builder(v, training, p + 'resnet_model/', updates, rng)
  net = stem(b, image)
  net = block_layers(b, net, resnet_size, 0, merge_after)
  if end_points is not None:
    end_points['pool2'] = net
  c = net.shape[-1]
  blocks = []
  for name in sorted(GRASP_PARAM_NAMES):
    off, size = GRASP_PARAM_NAMES[name]
    ensure(p + name + '/weights', (size, 256))
    ensure(p + name + '/biases', (256,), const=0)
    blocks.append(tf_ops.dense(grasp_params[:, off:off + size], v[p + name + '/weights'], v[p + name + '/biases'],
                               fp32_path=True))
  ctx = tf_ops.relu(head_bn(tf_ops._store(sum(blocks)), p + 'BatchNorm_1', scale=False))
  ensure(p + 'fcgrasp2/weights', (256, c))
  ctx = tf_ops.relu(head_bn(tf_ops.dense(ctx, v[p + 'fcgrasp2/weights']), p + 'fcgrasp2/BatchNorm'))
  if tile:
    net = net.repeat_interleave(a, dim=0)
  net = tf_ops._store(net + ctx.reshape(-1, 1, 1, c))
  net = block_layers(b, net, resnet_size, merge_after, 4)
  net = tf_ops._store(b.bn(net, relu=True).mean((1, 2)))
  k = net.shape[-1]
  for l in range(2):
    ensure(p + 'fc%d/weights' % l, (k, 64))
    net = tf_ops.relu(head_bn(tf_ops.dense(net, v[p + 'fc%d/weights' % l]), p + 'fc%d/BatchNorm' % l))
    k = 64
  ensure(p + 'logit/weights', (64, 1))
  ensure(p + 'logit/biases', (1,), const=0)
  logits = tf_ops.dense(net, v[p + 'logit/weights'], v[p + 'logit/biases'], fp32_path=True)
  if end_points is not None:
    end_points['logits'] = logits
    pred = torch.sigmoid(logits)
    end_points['predictions'] = pred.reshape(-1, a) if tile else pred
  return logits
