"""Grasp2Vec embedding towers (research/grasp2vec/networks.py:24-42, resnet.py:537-558)."""
from tensor2robot_b200 import nn
from tensor2robot_b200.layers import film_resnet_model as resnet_lib
from tensor2robot_b200.layers import resnet


def get_resnet50_spatial(images, is_training):
  """ResNet-50 v2 with the last block layer cut off; returns the pre-pooling `block_layer3` map
  (research/grasp2vec/resnet.py:537-558).  The head variables (final BN, 1001-way dense) are created
  like in the reference, so variable names / checkpoints line up."""
  model = resnet_lib.Model(
      resnet_size=50, bottleneck=True, num_classes=1001, num_filters=64, kernel_size=7, conv_stride=2,
      first_pool_size=3, first_pool_stride=2, block_sizes=[3, 4, 6], block_strides=[1, 2, 2],
      weight_decay=None, resnet_version=resnet_lib.DEFAULT_VERSION, data_format='channels_last')
  model(images, is_training)
  return resnet.resnet_endpoints(model)['block_layer3']


def Embedding(image, mode, params=None, scope='scene'):  # pylint: disable=invalid-name
  """(summed embedding [B, 1024] bf16, embedding map [B, h, w, 1024] bf16): ReLU of the truncated
  ResNet-50 map and its spatial mean (networks.py:24-42)."""
  del params
  is_training = mode == 'train'
  with nn.variable_scope(scope):
    scene = nn.relu(get_resnet50_spatial(image, is_training))
    summed_scene = nn.global_mean(scene)
  return summed_scene, scene
