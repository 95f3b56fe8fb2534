"""Spatial-softmax image tower and pose head of the pose_env / VRGripper models (layers/vision_layers.py:30-158,
277-350).  The layers have 32 channels on 64x64 frames - below one tensor-core tile - and run on the fp32 CUDA-core
kernels of csrc/vision_small.cu (`nn.conv2d_f32`, `nn.layer_norm`, `nn.spatial_softmax`); slim's arg_scope defaults
are spelled out per layer.  Not built: the batch-norm normaliser variant, FiLM conditioning (`film_output_params`,
`BuildFILMParams`) and the high-resolution tower."""
import torch

from tensor2robot_b200 import nn

LAYER_NORM = 'layer_norm'


def _check_normalizer(normalizer_fn):
  if normalizer_fn not in (LAYER_NORM, nn.layer_norm):
    raise NotImplementedError('only slim.layer_norm (the default) is built as the normaliser of vision_layers')


def BuildImagesToFeaturesModel(images, filter_size=3, num_blocks=5, num_output_maps=32, is_training=False,  # pylint: disable=invalid-name
                               normalizer_fn=LAYER_NORM, normalizer_params=None, weight_regularization=0.00001,
                               film_output_params=None, use_spatial_softmax=True):
  """images: fp32 [B, H, W, 3] in [0, 1].  num_blocks x (3x3 VALID conv 32, stride 2,2,1,1,..., LayerNorm, ReLU),
  a 1x1 conv (+LayerNorm, ReLU) to num_output_maps and the spatial softmax.  Returns (expected feature points
  [B, 2*num_output_maps] or the feature maps, {'softmax': heat map} or {}).  slim gives normalised convolutions no
  bias; the l2 regulariser only registers a collection entry that T2R models never add to the loss."""
  del is_training, normalizer_params, weight_regularization
  _check_normalizer(normalizer_fn)
  if film_output_params is not None:
    raise NotImplementedError('FiLM conditioning of the spatial-softmax tower is not built')
  net = images
  for i in range(num_blocks):
    scope = 'conv{:d}'.format(i + 2)
    net = nn.conv2d_f32(net, 32, filter_size, stride=2 if i < 2 else 1, padding='VALID', use_bias=False, scope=scope)
    net = nn.layer_norm(net, scope=scope + '/LayerNorm', relu=True)
  net = nn.conv2d_f32(net, num_output_maps, 1, stride=1, padding='VALID', use_bias=False, scope='final_conv_1x1')
  net = nn.layer_norm(net, scope='final_conv_1x1/LayerNorm', relu=True)
  if use_spatial_softmax:
    points, softmax = nn.spatial_softmax(net, return_softmax=True)
    return points, {'softmax': softmax}
  return net, {}


def BuildImageFeaturesToPoseModel(expected_feature_points, num_outputs, aux_input=None, aux_output_dim=0,  # pylint: disable=invalid-name
                                  hidden_dim=100, num_layers=2, is_training=True, normalizer_fn=LAYER_NORM,
                                  bias_transform_size=10):
  """Feature points [B, 2N] (+ aux_input) -> concat a learned bias-transform vector -> num_layers x
  (FC hidden_dim, LayerNorm, ReLU) -> FC num_outputs.  Returns (pose [B, num_outputs], aux output or None)."""
  del is_training
  _check_normalizer(normalizer_fn)
  net = expected_feature_points
  if aux_input is not None:
    net = torch.cat([net, aux_input.to(net.dtype)], 1)
  init = nn.truncated_normal(0.01)
  if bias_transform_size > 0:
    net = nn.bias_transform(net, bias_transform_size, scope='BiasAdd', initializer=0.01)
  for layer_index in range(num_layers):
    scope = 'pose_fc{:d}'.format(layer_index)
    net = nn.dense_f32(net, hidden_dim, scope=scope, bias_rows=0, initializer=init, regularize=False)
    net = nn.layer_norm(net, scope=scope + '/LayerNorm', relu=True)
  if num_outputs:
    net = nn.dense_f32(net, num_outputs, scope='pose_fc{:d}'.format(num_layers), initializer=init, regularize=False,
                       bias_initializer=0.01)
  if aux_output_dim > 0:
    aux = nn.dense_f32(expected_feature_points, aux_output_dim, scope='pose_fc_aux', initializer=init,
                       regularize=False, bias_initializer=0.01)
    return net, aux
  return net, None
