"""ORACLE (test infrastructure, never imported by the product): the reference's Grasping44 critic
restated on torch CPU fp32, statement by statement.

Follows research/qtopt/networks.py:343-615 (Grasping44FlexibleGraspParams.model) with the
grasp_param_names of :618-740 and the training objective of research/qtopt/t2r_models.py:229-239
(tf.losses.log_loss + regularisation losses).  PARITY UNPINNED: the reference has no golden values
for this network (its tests only check that it trains); this restatement is the
reference of record for the engine's parity tests.
"""
import collections

import numpy as np
import torch

from oracle import tf_ops

TOP_SCOPE = 'Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom'
GRASP_PARAM_NAMES = collections.OrderedDict([
    ('fcgrasp_wv', (0, 3)),
    ('fcgrasp_vr', (3, 2)),
    ('fcgrasp_gripper_close', (5, 1)),
    ('fcgrasp_gripper_open', (6, 1)),
    ('fcgrasp_terminate_episode', (7, 1)),
    ('fcgrasp_gripper_closed', (8, 1)),
    ('fcgrasp_height_to_bottom', (9, 1)),
])
BN_DECAY, BN_EPS, L2 = 0.9997, 0.001, 0.00007
NUM_CONVS = [6, 6, 3]


def init_variables(seed=0, scope=TOP_SCOPE):
  """Variables in TF names/layouts, truncated_normal(0.01) weights (networks.py:431-433)."""
  rng = np.random.RandomState(seed)

  def tn(shape, std=0.01):
    a = rng.normal(0, std, size=shape)
    a = np.clip(a, -2 * std, 2 * std)
    return a.astype(np.float32)

  v = collections.OrderedDict()

  def bn(name, c, scale=True):
    v[name + '/beta'] = np.zeros(c, np.float32)
    if scale:
      v[name + '/gamma'] = np.ones(c, np.float32)
    v[name + '/moving_mean'] = np.zeros(c, np.float32)
    v[name + '/moving_variance'] = np.ones(c, np.float32)

  p = scope + '/'
  v[p + 'conv1_1/weights'] = tn((6, 6, 3, 64))
  v[p + 'conv1_1/biases'] = np.zeros(64, np.float32)
  bn(p + 'BatchNorm', 64, scale=False)
  for l in range(2, 8):
    v[p + 'conv%d/weights' % l] = tn((5, 5, 64, 64))
    bn(p + 'conv%d/BatchNorm' % l, 64)
  for name in sorted(GRASP_PARAM_NAMES):
    v[p + name + '/weights'] = tn((GRASP_PARAM_NAMES[name][1], 256))
    v[p + name + '/biases'] = np.zeros(256, np.float32)
  bn(p + 'BatchNorm_1', 256, scale=False)
  v[p + 'fcgrasp2/weights'] = tn((256, 64))
  bn(p + 'fcgrasp2/BatchNorm', 64)
  for l in range(8, 17):
    v[p + 'conv%d/weights' % l] = tn((3, 3, 64, 64))
    bn(p + 'conv%d/BatchNorm' % l, 64)
  v[p + 'fc0/weights'] = tn((4096, 64))
  bn(p + 'fc0/BatchNorm', 64)
  v[p + 'fc1/weights'] = tn((64, 64))
  bn(p + 'fc1/BatchNorm', 64)
  v[p + 'logit/weights'] = tn((64, 1))
  v[p + 'logit/biases'] = np.zeros(1, np.float32)
  return v


def to_torch(variables, requires_grad=True):
  out = collections.OrderedDict()
  for k, a in variables.items():
    t = torch.tensor(np.asarray(a, np.float32))
    moving = k.endswith('moving_mean') or k.endswith('moving_variance')
    t.requires_grad_(requires_grad and not moving)
    out[k] = t
  return out


def regularized_names(variables):
  return [k for k in variables if k.endswith('/weights')]


def model(variables, image, grasp_params, is_training, scope=TOP_SCOPE, updates=None, end_points=None,
          goal_spatial=None, goal_vector=None):
  """image: [B,H,W,3] float32 in [0,1]; grasp_params: [B,10] or [B,A,10].  Returns logits [M,1].
  goal_spatial [G,h,w,C] / goal_vector [G,D]: what goal_spatial_fn() / goal_vector_fn() return (networks.py:548-561)."""
  v, p = variables, scope + '/'
  ep = end_points if end_points is not None else {}
  tile = grasp_params.dim() == 3
  a = grasp_params.shape[1] if tile else 1
  if tile:                                                         # networks.py:410-420
    grasp_params = grasp_params.reshape(-1, grasp_params.shape[2])

  def bn(x, name, scale=True):
    return tf_ops.batch_norm(x, v, p + name, is_training, BN_DECAY, BN_EPS, scale, updates)

  def conv_bn_relu(x, name, padding='SAME'):
    return tf_ops.relu(bn(tf_ops.conv2d(x, v[p + name + '/weights'], 1, padding), name + '/BatchNorm'))

  net = tf_ops.conv2d(image, v[p + 'conv1_1/weights'], 2, 'SAME', v[p + 'conv1_1/biases'])   # :443-450
  net = tf_ops.relu(bn(net, 'BatchNorm', scale=False))                                        # :459
  net = tf_ops.max_pool(net, 3, 3)                                                            # :460 pool1
  for l in range(2, 2 + NUM_CONVS[0]):                                                        # :462-464
    net = conv_bn_relu(net, 'conv%d' % l)
  net = tf_ops.max_pool(net, 3, 3)                                                            # :465 pool2
  ep['pool2'] = net
  blocks = []
  for name in sorted(GRASP_PARAM_NAMES):                                                      # :477-499
    off, size = GRASP_PARAM_NAMES[name]
    blocks.append(tf_ops.dense(grasp_params[:, off:off + size], v[p + name + '/weights'], v[p + name + '/biases'],
                               fp32_path=True))
  fcgrasp = tf_ops._store(sum(blocks))                                                        # :501 add_n
  fcgrasp = tf_ops.relu(bn(fcgrasp, 'BatchNorm_1', scale=False))                               # :511
  fcgrasp = tf_ops.relu(bn(tf_ops.dense(fcgrasp, v[p + 'fcgrasp2/weights']), 'fcgrasp2/BatchNorm'))  # :512
  ep['fcgrasp'] = fcgrasp
  context = fcgrasp.reshape(-1, 1, 1, 64)
  if tile:                                                                                    # :520-521 tile_batch
    net = net.repeat_interleave(a, dim=0)
  net = tf_ops._store(net + context)                                                          # :522
  ep['vsum'] = net
  for l in range(8, 14):                                                                      # :527-530
    net = conv_bn_relu(net, 'conv%d' % l)
  net = tf_ops.max_pool(net, 2, 2)                                                            # :532 pool3
  for l in range(14, 17):                                                                     # :536-539
    net = conv_bn_relu(net, 'conv%d' % l, 'VALID')
  ep['final_conv'] = net
  batch = net.shape[0]
  if goal_spatial is not None:                                                                # :548-553 tf.tile + concat
    net = torch.cat([net, tf_ops._store(goal_spatial).repeat(batch // goal_spatial.shape[0], 1, 1, 1)], dim=3)
  net = net.reshape(net.shape[0], -1)                                                         # :554 flatten
  if goal_vector is not None:                                                                 # :558-561
    net = torch.cat([net, tf_ops._store(goal_vector).repeat(batch // goal_vector.shape[0], 1)], dim=1)
  for l in range(2):                                                                          # :562-563
    net = tf_ops.relu(bn(tf_ops.dense(net, v[p + 'fc%d/weights' % l]), 'fc%d/BatchNorm' % l))
  logits = tf_ops.dense(net, v[p + 'logit/weights'], v[p + 'logit/biases'], fp32_path=True)   # :568-574
  ep['logits'] = logits
  pred = torch.sigmoid(logits)                                                                # :579
  ep['predictions'] = pred.reshape(-1, a) if tile else pred
  return logits


def train_loss(variables, image, grasp_params, reward, scope=TOP_SCOPE, updates=None):
  """t2r_models.py:229-239: log_loss(reward, sigmoid(logits)) + sum of l2 regularisers."""
  logits = model(variables, image, grasp_params, True, scope, updates)
  loss = tf_ops.log_loss(reward, torch.sigmoid(logits))
  reg = sum(tf_ops.l2_regularizer(L2, variables[k]) for k in regularized_names(variables))
  return loss + reg, loss, torch.sigmoid(logits)
