"""QT-Opt critic networks ("Grasping44" family) on the B200 engine.

Mirrors the public surface of the reference's research/qtopt/networks.py (class names, ctor
arguments, `model(...)` signature, variable names, end_points keys) with the graph re-expressed
over tensor2robot_b200.nn - i.e. tcgen05 implicit-GEMM convolutions, fused BN+ReLU kernels and
the never-materialised action tiling - instead of slim/TensorFlow ops.

Network definition followed: research/qtopt/networks.py:343-615 (Grasping44FlexibleGraspParams)
and :618-740 (the E2E open/close/terminate subclass).
"""
import torch

from tensor2robot_b200 import nn

NUM_LAYERS = 19
BATCH_SIZE = 64
NUM_SAMPLES = 100


class GraspingModel(object):
  """Base class of the grasping critics (research/qtopt/networks.py:38-60)."""

  def __init__(self, batch_norm_decay=0.9997, batch_norm_epsilon=0.001, l2_regularization=0.00007):
    self._batch_norm_decay = batch_norm_decay
    self._batch_norm_epsilon = batch_norm_epsilon
    self._l2_regularization = l2_regularization

  @property
  def l2_regularization(self):
    return self._l2_regularization

  @property
  def grasp_model_input_keys(self):
    return ['world_vector', 'vertical_rotation']

  def create_grasp_params_input(self, model_input, concat_axis=1):
    """Concatenates the action features named by grasp_model_input_keys (networks.py:60-73)."""
    return torch.cat([model_input[k].to(torch.float32) for k in self.grasp_model_input_keys], concat_axis)


class Grasping44FlexibleGraspParams(GraspingModel):
  """Grasping44 with a flat grasp_params vector; optional action batch (CEM megabatch)."""

  def __init__(self, action_batch_size=None, also_tile_batch_in_training=False, create_var_scope=True,
               **kwargs):
    super(Grasping44FlexibleGraspParams, self).__init__(**kwargs)
    self._action_batch_size = action_batch_size
    self._also_tile_batch_in_training = also_tile_batch_in_training
    self._create_var_scope = create_var_scope
    self.activation_layers = []
    self.num_convs = [6, 6, 3]
    self.hid_layers = 2

  # -- helpers --------------------------------------------------------------------------------
  def _conv_bn_relu(self, net, k, scope, is_training, padding='SAME'):
    """slim.conv2d with normalizer_fn=slim.batch_norm, activation relu (no bias)."""
    init = nn.truncated_normal(0.01)
    net = nn.conv2d(net, 64, k, 1, padding, use_bias=False, scope=scope, initializer=init,
                    defer_for_bn=not is_training)
    return nn.batch_norm(net, is_training, scope=scope + '/BatchNorm', scale=True, relu=True,
                         momentum=self._batch_norm_decay, eps=self._batch_norm_epsilon)

  def _fc_bn_relu(self, net, units, scope, is_training):
    net = nn.dense(net, units, scope=scope, initializer=nn.truncated_normal(0.01))
    return nn.batch_norm(net, is_training, scope=scope + '/BatchNorm', scale=True, relu=True,
                         momentum=self._batch_norm_decay, eps=self._batch_norm_epsilon)

  def image_tower(self, grasp_image, is_training, end_points=None):
    """conv1_1 .. pool2: the part shared by every action sample of a state (networks.py:443-467)."""
    init = nn.truncated_normal(0.01)
    net = nn.conv2d(grasp_image, 64, 6, 2, 'SAME', use_bias=True, scope='conv1_1', initializer=init)
    # stand-alone batch norm: scale=False (networks.py:451-461)
    net = nn.batch_norm(net, is_training, scope='BatchNorm', scale=False, relu=True,
                        momentum=self._batch_norm_decay, eps=self._batch_norm_epsilon)
    net = nn.max_pool2d(net, 3, 3, 'SAME')
    for l in range(2, 2 + self.num_convs[0]):
      net = self._conv_bn_relu(net, 5, 'conv%d' % l, is_training)
    net = nn.max_pool2d(net, 3, 3, 'SAME')
    if end_points is not None:
      end_points['pool2'] = net
    return net

  def action_context(self, grasp_params, grasp_param_names, is_training, end_points=None):
    """fcgrasp blocks -> add_n -> BN(scale=False)+ReLU -> fcgrasp2 (networks.py:469-512)."""
    if grasp_param_names is None:
      blocks = [('fcgrasp', 0, grasp_params.shape[1])]
    else:
      blocks = [(name, grasp_param_names[name][0], grasp_param_names[name][1])
                for name in sorted(grasp_param_names)]
    vs = nn.current_store()
    # All blocks share the 256-wide output, so their sum is one GEMM over the concatenated
    # input; block weights are row ranges of W, block biases rows of the bias matrix.
    order = sorted(range(len(blocks)), key=lambda i: blocks[i][1])
    if [blocks[i][1] for i in order] != [sum(blocks[j][2] for j in order[:n]) for n in range(len(order))]:
      raise ValueError('grasp_param_names must tile grasp_params without gaps')
    fc = nn.dense_f32(grasp_params, 256, scope='fcgrasp_blocks', bias_rows=len(blocks),
                      initializer=nn.truncated_normal(0.01))
    prefix = vs.full_name('')
    wv = vs.vars[vs.full_name('fcgrasp_blocks/weights')]
    bv = vs.vars[vs.full_name('fcgrasp_blocks/biases')]
    wv.tf_parts = [(prefix + name + '/weights', off, off + size, False) for name, off, size in blocks]
    bv.tf_parts = [(prefix + blocks[i][0] + '/biases', r, r + 1, True) for r, i in enumerate(order)]
    fc = nn.batch_norm(nn.to_bf16(fc), is_training, scope='BatchNorm_1', scale=False, relu=True,
                       momentum=self._batch_norm_decay, eps=self._batch_norm_epsilon)
    fc = self._fc_bn_relu(fc, 64, 'fcgrasp2', is_training)
    if end_points is not None:
      end_points['fcgrasp'] = fc
    return fc

  def q_head(self, net, is_training, num_classes, end_points, goal_spatial_fn=None, goal_vector_fn=None):
    """conv8.. -> pool3 -> conv14..16 (VALID) -> [goal merge] -> fc0, fc1 -> logit (networks.py:524-573)."""
    first = 2 + sum(self.num_convs[:1])
    for l in range(first, 2 + sum(self.num_convs[:2])):
      net = self._conv_bn_relu(net, 3, 'conv%d' % l, is_training)
    net = nn.max_pool2d(net, 2, 2, 'SAME')
    for l in range(2 + sum(self.num_convs[:2]), 2 + sum(self.num_convs[:3])):
      net = self._conv_bn_relu(net, 3, 'conv%d' % l, is_training, padding='VALID')
    end_points['final_conv'] = net
    batch = net.shape[0]
    if goal_spatial_fn is not None:
      # a goal feature map concatenated on the channel axis, tf.tile'd up to the (CEM-tiled) batch (networks.py:548-553)
      goal = nn.to_bf16(goal_spatial_fn())
      net = torch.cat([net, goal.repeat(batch // goal.shape[0], 1, 1, 1)], dim=3)
    net = net.reshape(batch, -1)  # slim.flatten: NHWC order
    if goal_vector_fn is not None:
      goal = nn.to_bf16(goal_vector_fn())                                             # networks.py:558-561
      net = torch.cat([net, goal.repeat(batch // goal.shape[0], 1)], dim=1)
    for l in range(self.hid_layers):
      if net.shape[1] % 64:
        # a goal merge can leave an inner dimension that the tensor-core path does not take: the fp32 FC kernel keeps
        # the reference's [K, 64] weight shape
        net = nn.to_bf16(nn.dense_f32(nn.to_f32(net), 64, scope='fc%d' % l, bias_rows=0,
                                      initializer=nn.truncated_normal(0.01)))
        net = nn.batch_norm(net, is_training, scope='fc%d/BatchNorm' % l, scale=True, relu=True,
                            momentum=self._batch_norm_decay, eps=self._batch_norm_epsilon)
      else:
        net = self._fc_bn_relu(net, 64, 'fc%d' % l, is_training)
    name = 'logit' if num_classes == 1 else 'logit_%d' % num_classes
    logits = nn.dense_f32(nn.to_f32(net), num_classes, scope=name, initializer=nn.truncated_normal(0.01))
    return logits

  # -- the graph ------------------------------------------------------------------------------
  def model(self, images, grasp_params, num_classes=1, is_training=False, softmax=False, restore=True,
            grasp_param_names=None, goal_spatial_fn=None, goal_vector_fn=None, scope=None, reuse=None,
            staged_features=None, **kwargs):
    """Builds/runs the critic.  Same contract as the reference's `model` plus `staged_features`:
    an already computed `pool2` map (CEM evaluates many action batches against one state tower,
    SURVEY 3.2)."""
    del kwargs, reuse
    if not restore:
      raise ValueError("This model doesn't yet support restore=False")
    if softmax:
      raise NotImplementedError('softmax head is not used by the QT-Opt critic')
    end_points = {}
    tile_batch = grasp_params.dim() == 3
    if tile_batch:
      a = grasp_params.shape[1]
      if self._action_batch_size is not None and a != self._action_batch_size:
        raise ValueError('grasp_params action dim %d != action_batch_size %d' % (a, self._action_batch_size))
      grasp_params = grasp_params.reshape(-1, grasp_params.shape[2])
    else:
      a = 1
    if scope is None:
      scope = self.__class__.__name__
    with nn.variable_scope(scope if self._create_var_scope else ''):
      if staged_features is not None:
        net = staged_features
        end_points['pool2'] = net
      else:
        _, grasp_image = images
        net = self.image_tower(grasp_image, is_training, end_points)
      context = self.action_context(grasp_params.to(torch.float32).contiguous(), grasp_param_names,
                                    is_training, end_points)
      net = nn.add_context(net, context, a)   # tile_batch + tf.add, never materialised
      end_points['vsum'] = net
      logits = self.q_head(net, is_training, num_classes, end_points, goal_spatial_fn, goal_vector_fn)
    end_points['logits'] = logits
    predictions = nn.sigmoid(logits.detach())
    if tile_batch:
      predictions = predictions.reshape(-1, a) if num_classes == 1 else predictions.reshape(-1, a, num_classes)
    end_points['predictions'] = predictions
    return logits, end_points


class Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom(Grasping44FlexibleGraspParams):
  """Grasping44 controlling gripper open/close/terminate, with gripper status and height
  to the bin bottom as extra state (research/qtopt/networks.py:618-740)."""

  GRASP_PARAM_NAMES = {
      'fcgrasp_wv': (0, 3),
      'fcgrasp_vr': (3, 2),
      'fcgrasp_gripper_close': (5, 1),
      'fcgrasp_gripper_open': (6, 1),
      'fcgrasp_terminate_episode': (7, 1),
      'fcgrasp_gripper_closed': (8, 1),
      'fcgrasp_height_to_bottom': (9, 1),
  }

  @property
  def grasp_model_input_keys(self):
    return ['world_vector', 'vertical_rotation', 'close_gripper', 'open_gripper', 'terminate_episode',
            'gripper_closed', 'height_to_bottom']

  def __call__(self, images, grasp_params, num_classes=1, is_training=False, softmax=False, restore=True,
               scope=None, reuse=None, **kwargs):
    return self.model(images, grasp_params, num_classes, is_training, softmax, restore, scope, reuse,
                      **kwargs)

  def model(self, images, grasp_params, num_classes=1, is_training=False, softmax=False, restore=True,
            scope=None, reuse=None, **kwargs):
    return super(Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom, self).model(
        images, grasp_params, num_classes=num_classes, is_training=is_training, softmax=softmax,
        restore=restore, scope=scope, reuse=reuse, grasp_param_names=dict(self.GRASP_PARAM_NAMES),
        **kwargs)
