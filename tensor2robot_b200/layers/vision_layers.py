"""Spatial-softmax image tower and pose head of the pose_env / VRGripper models (layers/vision_layers.py:30-158,
277-350).  The layers have 32 channels on 64x64 frames - below one tensor-core tile - and run on the fp32 CUDA-core
kernels of csrc/vision_small.cu (`nn.conv2d_f32`, `nn.layer_norm`, `nn.spatial_softmax`); slim's arg_scope defaults
are spelled out per layer.  Both normalisers of the reference are built - slim.layer_norm (the default) and
slim.batch_norm (decay .99, epsilon 1e-4, scale only on the final 1x1 convolution, :72-86) - as is the FiLM
conditioning of the tower (`film_output_params`, `BuildFILMParams`, :100-141, 162-181).  Not built: the
high-resolution tower."""
import torch

from tensor2robot_b200 import nn

LAYER_NORM = 'layer_norm'
BATCH_NORM = 'batch_norm'
NUM_CHANNELS_PER_BLOCK = 32


def _check_normalizer(normalizer_fn):
  if normalizer_fn in (LAYER_NORM, nn.layer_norm):
    return LAYER_NORM
  if normalizer_fn in (BATCH_NORM, nn.batch_norm_f32):
    return BATCH_NORM
  raise ValueError('normalizer_fn must be slim.layer_norm (%r) or slim.batch_norm (%r)' % (LAYER_NORM, BATCH_NORM))


def _normalize(net, kind, scope, is_training, params, relu):
  if kind == LAYER_NORM:
    return nn.layer_norm(net, scope=scope + '/LayerNorm', relu=relu)
  return nn.batch_norm_f32(net, params.get('is_training', is_training), scope=scope + '/BatchNorm',
                           scale=params.get('scale', False), decay=params.get('decay', 0.99),
                           eps=params.get('epsilon', 0.0001), relu=relu)


def BuildFILMParams(embedding, film_output_size=2 * 5 * 32):  # pylint: disable=invalid-name
  """A linear layer from the conditioning embedding [N, E] to the FiLM parameters [N, film_output_size]
  (layers/vision_layers.py:162-181): 2 x (total channels of the conditioned convolutions)."""
  return nn.dense_f32(embedding.float(), film_output_size, scope='film', regularize=False)


def BuildImagesToFeaturesModel(images, filter_size=3, num_blocks=5, num_output_maps=32, is_training=False,  # pylint: disable=invalid-name
                               normalizer_fn=LAYER_NORM, normalizer_params=None, weight_regularization=0.00001,
                               film_output_params=None, use_spatial_softmax=True):
  """images: fp32 [B, H, W, 3] in [0, 1].  num_blocks x (3x3 VALID conv 32, stride 2,2,1,1,..., LayerNorm, ReLU),
  a 1x1 conv (+LayerNorm, ReLU) to num_output_maps and the spatial softmax.  Returns (expected feature points
  [B, 2*num_output_maps] or the feature maps, {'softmax': heat map} or {}).  slim gives normalised convolutions no
  bias; the l2 regulariser only registers a collection entry that T2R models never add to the loss."""
  del weight_regularization
  kind = _check_normalizer(normalizer_fn)
  params, final_params = dict(normalizer_params or {}), dict(normalizer_params or {})
  if normalizer_params is None and kind == BATCH_NORM:      # the reference's defaults (:72-86)
    params = {'is_training': is_training, 'decay': 0.99, 'scale': False, 'epsilon': 0.0001}
    final_params = dict(params, scale=True)
  gammas_betas = None
  if film_output_params is not None:
    # [N, 2 * num_blocks * 32]: all gammas, then all betas, one 32-wide slice per conditioned block; applied as
    # (1 + gamma) * h + beta right before the ReLU (:100-141)
    expected = 2 * num_blocks * NUM_CHANNELS_PER_BLOCK
    if film_output_params.dim() != 2:
      raise ValueError('FILM shape is %s but is expected to be 2-D' % str(list(film_output_params.shape)))
    if film_output_params.shape[-1] != expected:
      raise ValueError('FILM shape is %s but final dimension should be %d' % (str(list(film_output_params.shape)), expected))
    film = film_output_params.float()
    half = num_blocks * NUM_CHANNELS_PER_BLOCK
    c = NUM_CHANNELS_PER_BLOCK
    gammas_betas = [torch.cat([film[:, i * c:(i + 1) * c], film[:, half + i * c:half + (i + 1) * c]], 1)
                    for i in range(num_blocks)]
  net = images
  for i in range(num_blocks):
    scope = 'conv{:d}'.format(i + 2)
    net = nn.conv2d_f32(net, NUM_CHANNELS_PER_BLOCK, filter_size, stride=2 if i < 2 else 1, padding='VALID',
                        use_bias=False, scope=scope)
    if gammas_betas is None:
      net = _normalize(net, kind, scope, is_training, params, relu=True)
    else:                                                  # Conv -> norm -> FiLM -> ReLU
      net = _normalize(net, kind, scope, is_training, params, relu=False)
      net = nn.film_relu_f32(net, gammas_betas[i])
  net = nn.conv2d_f32(net, num_output_maps, 1, stride=1, padding='VALID', use_bias=False, scope='final_conv_1x1')
  net = _normalize(net, kind, 'final_conv_1x1', is_training, final_params, relu=True)
  if use_spatial_softmax:
    points, softmax = nn.spatial_softmax(net, return_softmax=True)
    return points, {'softmax': softmax}
  return net, {}


def BuildImageFeaturesToPoseModel(expected_feature_points, num_outputs, aux_input=None, aux_output_dim=0,  # pylint: disable=invalid-name
                                  hidden_dim=100, num_layers=2, is_training=True, normalizer_fn=LAYER_NORM,
                                  bias_transform_size=10):
  """Feature points [B, 2N] (+ aux_input) -> concat a learned bias-transform vector -> num_layers x
  (FC hidden_dim, LayerNorm, ReLU) -> FC num_outputs.  Returns (pose [B, num_outputs], aux output or None)."""
  kind = _check_normalizer(normalizer_fn)
  if kind == BATCH_NORM:
    raise ValueError('BuildImageFeaturesToPoseModel: normalizer_fn cannot be batch norm (layers/vision_layers.py:303)')
  bn_params = {}
  net = expected_feature_points
  if aux_input is not None:
    net = torch.cat([net, aux_input.to(net.dtype)], 1)
  init = nn.truncated_normal(0.01)
  if bias_transform_size > 0:
    net = nn.bias_transform(net, bias_transform_size, scope='BiasAdd', initializer=0.01)
  for layer_index in range(num_layers):
    scope = 'pose_fc{:d}'.format(layer_index)
    net = nn.dense_f32(net, hidden_dim, scope=scope, bias_rows=0, initializer=init, regularize=False)
    net = _normalize(net, kind, scope, is_training, bn_params, relu=True)
  if num_outputs:
    net = nn.dense_f32(net, num_outputs, scope='pose_fc{:d}'.format(num_layers), initializer=init, regularize=False,
                       bias_initializer=0.01)
  if aux_output_dim > 0:
    aux = nn.dense_f32(expected_feature_points, aux_output_dim, scope='pose_fc_aux', initializer=init,
                       regularize=False, bias_initializer=0.01)
    return net, aux
  return net, None
