"""ResNet (v2, pre-activation) with optional FiLM conditioning on the B200 engine.

Mirrors the public surface of the reference's layers/film_resnet_model.py (`Model`, its ctor
arguments, `__call__(inputs, training, film_generator_fn, film_generator_input)`, block functions,
variable names as tf.layers would auto-number them under 'resnet_model/') with the graph expressed
over tensor2robot_b200.nn: every convolution is a tcgen05 implicit GEMM, the residual add is fused
into the epilogue of the block's last convolution, BN+FiLM+ReLU is one HBM pass.

Reference: layers/film_resnet_model.py:39-57 (batch_norm: momentum .997, eps 1e-5),
:60-105 (fixed_padding / conv2d_fixed_padding), :108-115 (_apply_film), :166-223 and :283-340
(v2 blocks), :343-388 (block_layer), :525-629 (Model.__call__).
"""
import torch

from tensor2robot_b200 import nn

_BATCH_NORM_DECAY = 0.997
_BATCH_NORM_EPSILON = 1e-5
DEFAULT_VERSION = 2


class _Namer(object):
  """tf.layers auto-numbering inside one variable scope: conv2d, conv2d_1, ..."""

  def __init__(self):
    self.counts = {}

  def __call__(self, base):
    n = self.counts.get(base, 0)
    self.counts[base] = n + 1
    return base if n == 0 else '%s_%d' % (base, n)


def batch_norm(inputs, training, namer, relu=False, film=None, passthrough=False, defer=False):
  """tf.layers.batch_normalization(momentum=.997, eps=1e-5, fused=True) [+FiLM] [+ReLU].
  passthrough=True additionally returns `inputs` routed through the same autograd node (used for
  the identity shortcut, whose gradient the BN backward kernel then adds for free).
  defer=True: may return an nn.DeferredBN that the consuming convolution fuses with (training only)."""
  return nn.batch_norm(inputs, training, scope=namer('batch_normalization'), scale=True, relu=relu,
                       momentum=_BATCH_NORM_DECAY, eps=_BATCH_NORM_EPSILON, film=film, passthrough=passthrough,
                       defer=defer)


def _conv_spec(filters, kernel_size, strides, namer, weight_decay):
  return dict(filters=filters, kernel_size=kernel_size, stride=strides,
              padding='SAME' if strides == 1 else 'FIXED', scope=namer('conv2d'),
              regularize=weight_decay is not None, names=('kernel', 'bias'))


def _preact_and_first_conv(inputs, training, namer, projection_shortcut, filters, kernel_size, strides,
                           weight_decay):
  """BN+ReLU, then the (optional) projection shortcut and the block's first convolution.

  Identity shortcut: `inputs` is returned through the BN node (gradient fan-in fused into the BN
  backward kernel).  Projection shortcut: both convolutions read the pre-activation, so they run as
  one autograd node whose second data gradient accumulates into the first (nn.conv2d_pair)."""
  if projection_shortcut is None:
    if training and torch.is_grad_enabled():
      out = batch_norm(inputs, training, namer, relu=True, passthrough=True, defer=True)
      if isinstance(out, nn.DeferredBN):   # BN + ReLU + conv as one node; the shortcut comes out of it
        first = conv2d_fixed_padding(out, filters, kernel_size, strides, namer, weight_decay)
        return out.shortcut, first
      preact, shortcut = out
    else:
      preact, shortcut = batch_norm(inputs, training, namer, relu=True), inputs
    first = conv2d_fixed_padding(preact, filters, kernel_size, strides, namer, weight_decay,
                                 defer_for_bn=not training)
    return shortcut, first
  preact = batch_norm(inputs, training, namer, relu=True, defer=True)
  fused = getattr(projection_shortcut, 'fused_args', None)
  if fused is None or not training:
    if isinstance(preact, nn.DeferredBN):
      preact = preact.materialize()
    shortcut = projection_shortcut(preact)
    first = conv2d_fixed_padding(preact, filters, kernel_size, strides, namer, weight_decay,
                                 defer_for_bn=not training)
    return shortcut, first
  proj_filters, proj_strides = fused
  shortcut, first = nn.conv2d_pair(preact, _conv_spec(proj_filters, 1, proj_strides, namer, weight_decay),
                                   _conv_spec(filters, kernel_size, strides, namer, weight_decay))
  return shortcut, first


def conv2d_fixed_padding(inputs, filters, kernel_size, strides, namer, weight_decay=None, residual=None,
                         needs_dgrad=True, defer_for_bn=False):
  """Strided convs use explicit (k-1)//2 padding + VALID, others SAME (film_resnet_model.py:89-105).
  The kernel variable is named 'kernel' like tf.layers.conv2d; weight_decay only marks the
  variable as regularised (the l2 gradient is applied by the fused optimizer kernel)."""
  padding = 'SAME' if strides == 1 else 'FIXED'
  return nn.conv2d(inputs, filters, kernel_size, strides, padding, use_bias=False, scope=namer('conv2d'),
                   regularize=weight_decay is not None, residual=residual, needs_dgrad=needs_dgrad,
                   names=('kernel', 'bias'), defer_for_bn=defer_for_bn)


def _film_tensor(film_gamma_beta):
  return None if film_gamma_beta is None else nn.to_f32(film_gamma_beta).contiguous()


def _building_block_v2(inputs, filters, training, projection_shortcut, strides, namer, weight_decay,
                       film_gamma_beta=None):
  """BN-ReLU-conv3x3-BN-[FiLM]-ReLU-conv3x3 + shortcut (film_resnet_model.py:166-223)."""
  shortcut, inputs = _preact_and_first_conv(inputs, training, namer, projection_shortcut, filters, 3, strides,
                                            weight_decay)
  inputs = batch_norm(inputs, training, namer, relu=True, film=_film_tensor(film_gamma_beta), defer=True)
  return conv2d_fixed_padding(inputs, filters, 3, 1, namer, weight_decay, residual=shortcut)


def _bottleneck_block_v2(inputs, filters, training, projection_shortcut, strides, namer, weight_decay,
                         film_gamma_beta=None):
  """BN-ReLU-1x1-BN-ReLU-3x3(stride)-BN-[FiLM]-ReLU-1x1(4x) + shortcut (film_resnet_model.py:283-340)."""
  shortcut, inputs = _preact_and_first_conv(inputs, training, namer, projection_shortcut, filters, 1, 1,
                                            weight_decay)
  inputs = batch_norm(inputs, training, namer, relu=True, defer=True)
  inputs = conv2d_fixed_padding(inputs, filters, 3, strides, namer, weight_decay, defer_for_bn=not training)
  inputs = batch_norm(inputs, training, namer, relu=True, film=_film_tensor(film_gamma_beta), defer=True)
  return conv2d_fixed_padding(inputs, 4 * filters, 1, 1, namer, weight_decay, residual=shortcut)


def _v1_shortcut(inputs, training, projection_shortcut, namer):
  """v1: the projection shortcut is followed by its own batch norm (film_resnet_model.py:149-153)."""
  if projection_shortcut is None:
    return inputs
  return batch_norm(projection_shortcut(inputs), training, namer)


def _building_block_v1(inputs, filters, training, projection_shortcut, strides, namer, weight_decay,
                       film_gamma_beta=None):
  """conv3x3-BN-ReLU-conv3x3-BN-[FiLM] + shortcut, ReLU (film_resnet_model.py:121-168)."""
  shortcut = _v1_shortcut(inputs, training, projection_shortcut, namer)
  inputs = conv2d_fixed_padding(inputs, filters, 3, strides, namer, weight_decay, defer_for_bn=not training)
  inputs = batch_norm(inputs, training, namer, relu=True)
  inputs = conv2d_fixed_padding(inputs, filters, 3, 1, namer, weight_decay, defer_for_bn=not training)
  inputs = batch_norm(inputs, training, namer, film=_film_tensor(film_gamma_beta))
  return nn.add_relu(inputs, shortcut)


def _bottleneck_block_v1(inputs, filters, training, projection_shortcut, strides, namer, weight_decay,
                         film_gamma_beta=None):
  """1x1-BN-ReLU-3x3(stride)-BN-ReLU-1x1(4x)-BN-[FiLM] + shortcut, ReLU (film_resnet_model.py:220-276)."""
  shortcut = _v1_shortcut(inputs, training, projection_shortcut, namer)
  inputs = conv2d_fixed_padding(inputs, filters, 1, 1, namer, weight_decay, defer_for_bn=not training)
  inputs = batch_norm(inputs, training, namer, relu=True)
  inputs = conv2d_fixed_padding(inputs, filters, 3, strides, namer, weight_decay, defer_for_bn=not training)
  inputs = batch_norm(inputs, training, namer, relu=True)
  inputs = conv2d_fixed_padding(inputs, 4 * filters, 1, 1, namer, weight_decay, defer_for_bn=not training)
  inputs = batch_norm(inputs, training, namer, film=_film_tensor(film_gamma_beta))
  return nn.add_relu(inputs, shortcut)


def block_layer(inputs, filters, bottleneck, block_fn, blocks, strides, training, name, namer,
                weight_decay, film_gamma_betas):
  """One block layer; only the first block projects and strides (film_resnet_model.py:343-388)."""
  del name
  if blocks != len(film_gamma_betas):
    raise ValueError('film_gamma_betas has length {}, expected {}'.format(len(film_gamma_betas), blocks))
  filters_out = filters * 4 if bottleneck else filters

  def projection_shortcut(x):
    return conv2d_fixed_padding(x, filters_out, 1, strides, namer, weight_decay)
  projection_shortcut.fused_args = (filters_out, strides)

  inputs = block_fn(inputs, filters, training, projection_shortcut, strides, namer, weight_decay,
                    film_gamma_betas[0])
  for i in range(1, blocks):
    inputs = block_fn(inputs, filters, training, None, 1, namer, weight_decay, film_gamma_betas[i])
  return inputs


class Model(object):
  """ResNet builder (film_resnet_model.py:391-629)."""

  def __init__(self, resnet_size, bottleneck, num_classes, num_filters, kernel_size, conv_stride,
               first_pool_size, first_pool_stride, block_sizes, block_strides, weight_decay,
               resnet_version=DEFAULT_VERSION, data_format=None, dtype=torch.float32):
    if resnet_version not in (1, 2):
      raise ValueError('Resnet version should be 1 or 2. See README for citations.')
    if data_format not in (None, 'channels_last'):
      raise ValueError('the B200 engine is NHWC (channels_last) only')
    self.resnet_size = resnet_size
    self.resnet_version = resnet_version
    self.bottleneck = bottleneck
    if resnet_version == 1:
      self.block_fn = _bottleneck_block_v1 if bottleneck else _building_block_v1
    else:
      self.block_fn = _bottleneck_block_v2 if bottleneck else _building_block_v2
    self.data_format = 'channels_last'
    self.num_classes = num_classes
    self.num_filters = num_filters
    self.kernel_size = kernel_size
    self.conv_stride = conv_stride
    self.first_pool_size = first_pool_size
    self.first_pool_stride = first_pool_stride
    self.block_sizes = block_sizes
    self.block_strides = block_strides
    self.weight_decay = weight_decay
    self.dtype = dtype
    self.pre_activation = resnet_version == 2
    self.end_points = {}

  # The three stages are exposed separately so that a critic can merge the action context
  # between block layers (SURVEY A-15); __call__ chains them exactly like the reference.
  def stem(self, inputs, namer, training=False):
    inputs = conv2d_fixed_padding(inputs, self.num_filters, self.kernel_size, self.conv_stride, namer,
                                  self.weight_decay, needs_dgrad=False)
    self.end_points['initial_conv'] = inputs
    if self.resnet_version == 1:        # v2 leaves BN + ReLU to the first block's pre-activation (:565-571)
      inputs = batch_norm(inputs, training, namer, relu=True)
    if self.first_pool_size:
      inputs = nn.max_pool2d(inputs, self.first_pool_size, self.first_pool_stride, 'SAME')
      self.end_points['initial_max_pool'] = inputs
    return inputs

  def block_layers(self, inputs, training, namer, film_gamma_betas, first=0, last=None):
    last = len(self.block_sizes) if last is None else last
    for i in range(first, last):
      num_blocks = self.block_sizes[i]
      num_filters = self.num_filters * (2**i)
      if film_gamma_betas[i] is None:
        continue
      if len(film_gamma_betas[i]) != num_blocks:
        raise ValueError('Got {} FiLM vectors for block {}, expected {}'.format(
            len(film_gamma_betas[i]), i, num_blocks))
      for film_gamma_beta in film_gamma_betas[i]:
        if film_gamma_beta is None:
          continue
        film_shape = list(film_gamma_beta.shape)
        if len(film_shape) != 2:
          raise ValueError('FILM shape is %s but is expected to be 2-D' % str(film_shape))
        if film_shape[-1] != 2 * num_filters:
          raise ValueError('FILM shape is %s but final dimension should be %d' % (str(film_shape), 2 * num_filters))
      inputs = block_layer(inputs, num_filters, self.bottleneck, self.block_fn, num_blocks,
                           self.block_strides[i], training, 'block_layer{}'.format(i + 1), namer,
                           self.weight_decay, film_gamma_betas[i])
      self.end_points['block_layer{}'.format(i + 1)] = inputs
    return inputs

  def head(self, inputs, training, namer, dense=True):
    if self.pre_activation:
      inputs = batch_norm(inputs, training, namer, relu=True)
    self.end_points['pre_final_pool'] = inputs
    inputs = nn.global_mean(inputs)
    self.end_points['final_reduce_mean'] = inputs
    if dense:
      k = inputs.shape[1]
      if k % 64 == 0 and self.num_classes % 64 == 0:
        inputs = nn.dense(inputs, self.num_classes, scope=namer('dense'), use_bias=True,
                          regularize=False, names=('kernel', 'bias'))
      else:
        inputs = nn.dense_f32(nn.to_f32(inputs), self.num_classes, scope=namer('dense'), regularize=False,
                              names=('kernel', 'bias'))
      self.end_points['final_dense'] = inputs
    return inputs

  def film_params(self, film_generator_fn, film_generator_input):
    if film_generator_input is not None and film_generator_fn is None:
      raise ValueError('film_generator_input is provided but film_generator_fn is not specified.')
    if film_generator_fn:
      filter_sizes = [self.num_filters * (2**i) for i in range(len(self.block_sizes))]
      return film_generator_fn(film_generator_input, self.block_sizes, filter_sizes)
    return [[None] * n for n in self.block_sizes]

  def __call__(self, inputs, training, film_generator_fn=None, film_generator_input=None):
    """[N,H,W,C] bf16 images -> [N, num_classes] logits."""
    film_gamma_betas = self.film_params(film_generator_fn, film_generator_input)
    self.end_points = {}
    namer = _Namer()
    with nn.variable_scope('resnet_model'):
      inputs = self.stem(inputs, namer, training)
      inputs = self.block_layers(inputs, training, namer, film_gamma_betas)
      return self.head(inputs, training, namer)
