"""TEST INFRASTRUCTURE ONLY - CPU restatement (torch float64, autograd-capable) of the pose_env networks:
layers/vision_layers.py:30-158 (BuildImagesToFeaturesModel), :277-350 (BuildImageFeaturesToPoseModel),
research/pose_env/pose_env_models.py:118-181 (MC critic) with slim's defaults
(research/dql_grasping_lib/tf_modules.py:25-44; slim.layer_norm: moments over all non-batch axes, eps 1e-12,
per-channel gamma / beta; slim.conv2d / fully_connected drop the bias when a normaliser is given).
Parity unpinned against TensorFlow itself (TF is not installed here): the restatement follows the reference code
and slim's documented semantics; the spatial softmax part is pinned by oracle/spatial_softmax.py's tests.
Weights are passed as {TF variable name: float64 tensor} in TF layouts (conv HWIO, fc [in, out])."""
import torch
import torch.nn.functional as F


def layer_norm(x, gamma, beta, eps=1e-12):
  dims = tuple(range(1, x.dim()))
  mean = x.mean(dims, keepdim=True)
  var = ((x - mean) ** 2).mean(dims, keepdim=True)
  return (x - mean) / torch.sqrt(var + eps) * gamma + beta


def conv_valid(x_nhwc, w_hwio, stride):
  y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), w_hwio.permute(3, 2, 0, 1), stride=stride)
  return y.permute(0, 2, 3, 1)


def spatial_softmax(net):
  """layers/spatial_softmax.py:45-88: the points come out interleaved (x_1, y_1, x_2, y_2, ...)."""
  b, h, w, c = net.shape
  logits = net.permute(0, 3, 1, 2).reshape(b * c, h * w)
  sm = torch.softmax(logits, 1)
  xs = torch.linspace(-1.0, 1.0, w, dtype=net.dtype).repeat(h)
  ys = torch.linspace(-1.0, 1.0, h, dtype=net.dtype).repeat_interleave(w)
  ex, ey = (sm * xs).sum(1, keepdim=True), (sm * ys).sum(1, keepdim=True)
  return torch.cat([ex, ey], 1).reshape(b, 2 * c)


def batch_norm(x, w, scope, training, scale, eps=1e-4, decay=0.99, updates=None):
  """slim.batch_norm as vision_layers.py:72-86 configures it: batch statistics (biased variance) in training, the
  moving averages otherwise; the moving variance receives the Bessel-corrected batch variance (fused batch norm)."""
  dims = tuple(range(x.dim() - 1))
  if training:
    mean = x.mean(dims)
    var = ((x - mean) ** 2).mean(dims)
    if updates is not None:
      n = x.numel() // x.shape[-1]
      updates[scope + '/moving_mean'] = w[scope + '/moving_mean'] * decay + mean.detach() * (1 - decay)
      updates[scope + '/moving_variance'] = w[scope + '/moving_variance'] * decay + var.detach() * n / (n - 1) * (1 - decay)
  else:
    mean, var = w[scope + '/moving_mean'], w[scope + '/moving_variance']
  y = (x - mean) / torch.sqrt(var + eps)
  if scale:
    y = y * w[scope + '/gamma']
  return y + w[scope + '/beta']


def images_to_features(images, w, prefix, num_blocks=5, film=None, normalizer='layer_norm', training=True, updates=None):
  """film: [N, 2 * num_blocks * 32] (all gammas, then all betas), applied as (1 + gamma) * h + beta before the ReLU
  (vision_layers.py:100-141).  normalizer 'batch_norm': scale only on the final 1x1 convolution (:72-86)."""
  net = images

  def norm(x, s, scale):
    if normalizer == 'layer_norm':
      return layer_norm(x, w[s + '/LayerNorm/gamma'], w[s + '/LayerNorm/beta'])
    return batch_norm(x, w, s + '/BatchNorm', training, scale, updates=updates)

  for i in range(num_blocks):
    s = '%s/conv%d' % (prefix, i + 2)
    net = conv_valid(net, w[s + '/weights'], 2 if i < 2 else 1)
    net = norm(net, s, False)
    if film is not None:
      half = num_blocks * 32
      gamma = 1.0 + film[:, i * 32:(i + 1) * 32][:, None, None, :]
      beta = film[:, half + i * 32:half + (i + 1) * 32][:, None, None, :]
      net = gamma * net + beta
    net = torch.relu(net)
  s = prefix + '/final_conv_1x1'
  net = conv_valid(net, w[s + '/weights'], 1)
  net = torch.relu(norm(net, s, True))
  return spatial_softmax(net)


def features_to_pose(points, w, prefix, num_layers=2):
  net = torch.cat([points, w[prefix + '/BiasAdd/biases'].reshape(1, -1).expand(points.shape[0], -1)], 1)
  for i in range(num_layers):
    s = '%s/pose_fc%d' % (prefix, i)
    net = net @ w[s + '/weights']
    net = torch.relu(layer_norm(net, w[s + '/LayerNorm/gamma'], w[s + '/LayerNorm/beta']))
  s = '%s/pose_fc%d' % (prefix, num_layers)
  return net @ w[s + '/weights'] + w[s + '/biases'].reshape(1, -1)


def regression_a_func(images, w, prefix='a_func'):
  points = images_to_features(images, w, prefix + '/state_features')
  return features_to_pose(points, w, prefix), points


def mc_critic_q(images, pose, w, prefix='q_func'):
  net = images
  for i in range(3):
    s = '%s/q_features/%s' % (prefix, 'Conv' if i == 0 else 'Conv_%d' % i)
    net = conv_valid(net, w[s + '/weights'], 2)
    net = torch.relu(layer_norm(net, w[s + '/LayerNorm/gamma'], w[s + '/LayerNorm/beta']))
  s = prefix + '/q_features/fully_connected'
  ctx = torch.relu(pose @ w[s + '/weights'] + w[s + '/biases'].reshape(1, -1))
  net = net.repeat(ctx.shape[0] // net.shape[0], 1, 1, 1) + ctx[:, None, None, :]     # tf.tile of the batch
  net = net.reshape(net.shape[0], -1)
  for i in (1, 2):
    s = '%s/Stack/fully_connected_%d' % (prefix, i)
    net = torch.relu(net @ w[s + '/weights'] + w[s + '/biases'].reshape(1, -1))
  s = prefix + '/fully_connected'
  return (net @ w[s + '/weights'] + w[s + '/biases'].reshape(1, -1)).squeeze(1)


def weighted_mse(labels, predictions, weights=1.0):
  """tf.losses.mean_squared_error, SUM_BY_NONZERO_WEIGHTS."""
  w = torch.broadcast_to(torch.as_tensor(weights, dtype=predictions.dtype), predictions.shape)
  n = (w != 0).sum().clamp(min=1)
  return ((predictions - labels) ** 2 * w).sum() / n
