"""resnet_model / linear_film_generator with the reference's signatures (layers/resnet.py:31-209)."""
from tensor2robot_b200 import nn
from tensor2robot_b200.layers import film_resnet_model as resnet_lib


def _get_block_sizes(resnet_size):
  """Blocks per block layer (layers/resnet.py:31-62)."""
  choices = {
      18: [2, 2, 2, 2],
      34: [3, 4, 6, 3],
      50: [3, 4, 6, 3],
      101: [3, 4, 23, 3],
      152: [3, 8, 36, 3],
      200: [3, 24, 36, 3]
  }
  try:
    return choices[resnet_size]
  except KeyError:
    raise ValueError('Could not find layers for selected Resnet size.\n'
                     'Size received: {}; sizes allowed: {}.'.format(resnet_size, list(choices.keys())))


def resnet_endpoints(model):
  """Intermediate activations by the reference's tensor names (layers/resnet.py:80-94)."""
  names = ['initial_conv', 'initial_max_pool', 'pre_final_pool', 'final_reduce_mean', 'final_dense']
  names += ['block_layer{}'.format(i + 1) for i in range(len(model.block_sizes))]
  return {n: model.end_points[n] for n in names if n in model.end_points}


def linear_film_generator(embedding, block_sizes, filter_sizes, enabled_block_layers=None):
  """One Linear per block layer emitting num_blocks*C*2 values, split per block
  (layers/resnet.py:98-143).  film_gamma_betas[i][j] is None or a [batch, 2*C] tensor."""
  if enabled_block_layers:
    if len(enabled_block_layers) != len(block_sizes):
      raise ValueError('Got {} bools for enabled_block_layers, expected {}'.format(
          len(enabled_block_layers), len(block_sizes)))
  film_gamma_betas = []
  tensor_core = embedding.shape[-1] % 64 == 0     # e.g. the 512-wide sentence embedding; one-hot task ids are not
  emb = nn.to_bf16(embedding) if tensor_core else nn.to_f32(embedding)
  for i, num_blocks in enumerate(block_sizes):
    if enabled_block_layers and not enabled_block_layers[i]:
      film_gamma_betas.append([None] * num_blocks)
      continue
    num_filters = filter_sizes[i]
    if tensor_core:
      out = nn.dense(emb, num_blocks * num_filters * 2, scope='film{}'.format(i), use_bias=True)
    else:
      out = nn.dense_f32(emb, num_blocks * num_filters * 2, scope='film{}'.format(i))
    film_gamma_betas.append(list(out.split(num_filters * 2, dim=-1)))
  return film_gamma_betas


def resnet_model(images, is_training, num_classes, resnet_size=50, weight_decay=None, kernel_size=7,
                 num_filters=64, return_intermediate_values=False, film_generator_fn=None,
                 film_generator_input=None, pretrain_checkpoint=None):
  """Runs the ResNet tower on NHWC bf16 `images`; returns logits or the end_points dict."""
  if pretrain_checkpoint:
    raise NotImplementedError('TF checkpoint warm start: load with VariableStore.import_tf instead')
  model = resnet_lib.Model(
      resnet_size=resnet_size,
      bottleneck=resnet_size >= 50,
      num_classes=num_classes,
      num_filters=num_filters,
      kernel_size=kernel_size,
      conv_stride=2,
      first_pool_size=3,
      first_pool_stride=2,
      block_sizes=_get_block_sizes(resnet_size),
      block_strides=[1, 2, 2, 2],
      resnet_version=resnet_lib.DEFAULT_VERSION,
      data_format='channels_last',
      weight_decay=weight_decay)
  final_dense = model(images, is_training, film_generator_fn, film_generator_input)
  if return_intermediate_values:
    return resnet_endpoints(model)
  return final_dense
