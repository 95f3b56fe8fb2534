"""QT-Opt Q-critic with a ResNet-50 (v2) vision tower: BASELINE.json config C2/C3.

This is NOT a reference symbol: the reference's QT-Opt critic is Grasping44 and
its ResNet-50 exists only as the generic tower layers/resnet.py.  The composition below is ours,
assembled from reference parts so that each part keeps a parity anchor:

  state tower   : layers/film_resnet_model.Model (ResNet-50 v2) stem + block layers 1..3
                  on the state image [B,472,472,3]            -> [B,30,30,1024]
  action context: the Grasping44 action branch (research/qtopt/networks.py:469-512):
                  fcgrasp blocks (10->256, summed) -> BN(scale=False)+ReLU -> FC(256->1024)+BN+ReLU
  merge         : tile_batch + broadcast add after block layer 3, mirroring `vsum` after `pool2`
                  (networks.py:513-522); in PREDICT the tower output is tiled A times implicitly
  post-merge    : block layer 4 on [B*A,...] -> final BN+ReLU -> mean over HW  -> [B*A,2048]
  head          : FC64+BN+ReLU, FC64+BN+ReLU, FC1 (bias) -> sigmoid  (networks.py:560-579)
"""
import torch

from tensor2robot_b200 import nn
from tensor2robot_b200.layers import film_resnet_model as resnet_lib
from tensor2robot_b200.layers import resnet as resnet_factory
from tensor2robot_b200.research.qtopt import networks


class ResNet50QCritic(networks.Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom):
  """Same action features / ctor surface as the Grasping44 E2E critic, ResNet-50 v2 tower."""

  def __init__(self, action_batch_size=None, also_tile_batch_in_training=False, resnet_size=50,
               merge_after_block_layer=3, **kwargs):
    super(ResNet50QCritic, self).__init__(action_batch_size=action_batch_size,
                                          also_tile_batch_in_training=also_tile_batch_in_training, **kwargs)
    self._resnet_size = resnet_size
    self._merge_after = merge_after_block_layer
    self._resnet = resnet_lib.Model(
        resnet_size=resnet_size, bottleneck=resnet_size >= 50, num_classes=64, num_filters=64,
        kernel_size=7, conv_stride=2, first_pool_size=3, first_pool_stride=2,
        block_sizes=resnet_factory._get_block_sizes(resnet_size), block_strides=[1, 2, 2, 2],
        weight_decay=self._l2_regularization, data_format='channels_last')

  def image_tower(self, grasp_image, is_training, end_points=None, namer=None):
    namer = namer or resnet_lib._Namer()
    films = self._resnet.film_params(None, None)
    with nn.variable_scope('resnet_model'):
      net = self._resnet.stem(grasp_image, namer, is_training)
      net = self._resnet.block_layers(net, is_training, namer, films, 0, self._merge_after)
    if end_points is not None:
      end_points['pool2'] = net  # the staged feature map (name kept from Grasping44)
    return net, namer

  def _context(self, grasp_params, channels, is_training, end_points):
    vs = nn.current_store()
    blocks = [(name, self.GRASP_PARAM_NAMES[name][0], self.GRASP_PARAM_NAMES[name][1])
              for name in sorted(self.GRASP_PARAM_NAMES)]
    order = sorted(range(len(blocks)), key=lambda i: blocks[i][1])
    fc = nn.dense_f32(grasp_params, 256, scope='fcgrasp_blocks', bias_rows=len(blocks),
                      initializer=nn.truncated_normal(0.01))
    prefix = vs.full_name('')
    wv = vs.vars[vs.full_name('fcgrasp_blocks/weights')]
    bv = vs.vars[vs.full_name('fcgrasp_blocks/biases')]
    wv.tf_parts = [(prefix + name + '/weights', off, off + size, False) for name, off, size in blocks]
    bv.tf_parts = [(prefix + blocks[i][0] + '/biases', r, r + 1, True) for r, i in enumerate(order)]
    fc = nn.batch_norm(nn.to_bf16(fc), is_training, scope='BatchNorm_1', scale=False, relu=True,
                       momentum=self._batch_norm_decay, eps=self._batch_norm_epsilon)
    fc = nn.dense(fc, channels, scope='fcgrasp2', initializer=nn.truncated_normal(0.01))
    fc = nn.batch_norm(fc, is_training, scope='fcgrasp2/BatchNorm', scale=True, relu=True,
                       momentum=self._batch_norm_decay, eps=self._batch_norm_epsilon)
    end_points['fcgrasp'] = fc
    return fc

  def model(self, images, grasp_params, num_classes=1, is_training=False, softmax=False, restore=True,
            scope=None, reuse=None, staged_features=None, **kwargs):
    del kwargs, reuse, restore
    if softmax or num_classes != 1:
      raise NotImplementedError('the Q-critic has a single sigmoid output')
    end_points = {}
    tile_batch = grasp_params.dim() == 3
    a = grasp_params.shape[1] if tile_batch else 1
    if tile_batch:
      grasp_params = grasp_params.reshape(-1, grasp_params.shape[2])
    scope = scope or self.__class__.__name__
    with nn.variable_scope(scope):
      if staged_features is not None:
        net, namer = staged_features
        namer = _clone_namer(namer)
      else:
        net, namer = self.image_tower(images[1], is_training, end_points)
        self.staged = (net, _clone_namer(namer))
      context = self._context(grasp_params.to(torch.float32).contiguous(), net.shape[-1], is_training,
                              end_points)
      net = nn.add_context(net, context, a, defer_for_bn=not is_training)
      if not isinstance(net, nn.DeferredContext):   # fused into the next batch norm in inference graphs
        end_points['vsum'] = net
      films = self._resnet.film_params(None, None)
      with nn.variable_scope('resnet_model'):
        net = self._resnet.block_layers(net, is_training, namer, films, self._merge_after, None)
        net = self._resnet.head(net, is_training, namer, dense=False)
      for l in range(self.hid_layers):
        net = self._fc_bn_relu(net, 64, 'fc%d' % l, is_training)
      logits = nn.dense_f32(nn.to_f32(net), 1, scope='logit', initializer=nn.truncated_normal(0.01))
    end_points['logits'] = logits
    predictions = nn.sigmoid(logits.detach())
    end_points['predictions'] = predictions.reshape(-1, a) if tile_batch else predictions
    return logits, end_points


def _clone_namer(namer):
  c = resnet_lib._Namer()
  c.counts = dict(namer.counts)
  return c
