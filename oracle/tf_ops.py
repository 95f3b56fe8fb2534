"""ORACLE (test infrastructure, never imported by the product): TensorFlow-1.x op semantics
restated on torch CPU fp32.

The reference delegates all arithmetic to TensorFlow / tf_slim, which are absent from
/root/reference and from this image, so each function states the published
TF 1.15 semantics of the op the reference calls and cites the reference call site.
PARITY UNPINNED for network numerics: the reference ships no golden values for them
(SURVEY 4, 8c); these restatements are the reference of record.

Variables are passed as a dict {tf_variable_name: torch tensor in TF layout}
(conv kernels HWIO, dense kernels [in,out]).
"""
import math

import torch
import torch.nn.functional as F


TRACE = None   # list collecting (op, output) for layer-by-layer parity debugging


# Storage-precision model.  None = the reference's fp32 everywhere.  torch.bfloat16 = round every
# stored activation (conv / dense output, BN(+ReLU) output, pooled map) to bf16 like the engine
# does, keeping fp32 arithmetic inside each op: the "bf16-storage" oracle the engine must match
# tightly, while the fp32 oracle measures what bf16 storage costs.
STORAGE_DTYPE = None


def _store(y):
  if STORAGE_DTYPE is None:
    return y
  # straight-through rounding so that autograd still works in the oracle
  return y + (y.detach().to(STORAGE_DTYPE).to(y.dtype) - y.detach())


def _trace(op, y):
  if op != 'bn':          # bn is stored after the ReLU that follows it (see relu())
    y = _store(y)
  if TRACE is not None:
    TRACE.append((op, y))
  return y


def relu(x):
  """tf.nn.relu; the engine fuses it into the BN kernel and stores the result."""
  return _store(torch.relu(x))


def same_pad(size, k, stride):
  """TF 'SAME': out = ceil(in/s); pad_total = max((out-1)*s + k - in, 0); before = total//2."""
  out = -(-size // stride)
  total = max((out - 1) * stride + k - size, 0)
  return out, total // 2, total - total // 2


def conv2d(x, w_hwio, stride=1, padding='SAME', bias=None, residual=None):
  """tf.nn.conv2d on NHWC input with an HWIO kernel (slim.conv2d: networks.py:443; tf.layers.conv2d:
  film_resnet_model.py:99-105).  Zero padding, extra pixel at the bottom/right for SAME."""
  kh, kw = w_hwio.shape[0], w_hwio.shape[1]
  xn = x.permute(0, 3, 1, 2)
  if padding == 'SAME':
    _, pt, pb = same_pad(x.shape[1], kh, stride)
    _, pl, pr = same_pad(x.shape[2], kw, stride)
    xn = F.pad(xn, (pl, pr, pt, pb))
  elif padding != 'VALID':
    raise ValueError(padding)
  y = F.conv2d(xn, _store(w_hwio).permute(3, 2, 0, 1), bias, stride=stride)
  y = y.permute(0, 2, 3, 1)
  if residual is not None:   # `inputs + shortcut` (film_resnet_model.py:340): the engine fuses it into the conv
    y = y + residual
  return _trace('conv', y)


def fixed_padding(x, kernel_size):
  """film_resnet_model.py:60-86."""
  total = kernel_size - 1
  beg = total // 2
  end = total - beg
  return F.pad(x, (0, 0, beg, end, beg, end))


def conv2d_fixed_padding(x, w_hwio, strides, residual=None):
  """film_resnet_model.py:89-105: explicit padding + VALID when strided, SAME otherwise."""
  if strides > 1:
    x = fixed_padding(x, w_hwio.shape[0])
  return conv2d(x, w_hwio, strides, 'SAME' if strides == 1 else 'VALID', residual=residual)


def max_pool(x, k, stride, padding='SAME'):
  """slim.max_pool2d / tf.layers.max_pooling2d: padding never wins (-inf)."""
  xn = x.permute(0, 3, 1, 2)
  if padding == 'SAME':
    _, pt, pb = same_pad(x.shape[1], k, stride)
    _, pl, pr = same_pad(x.shape[2], k, stride)
    xn = F.pad(xn, (pl, pr, pt, pb), value=float('-inf'))
  return _trace('pool', F.max_pool2d(xn, k, stride).permute(0, 2, 3, 1))


def batch_norm(x, variables, scope, training, decay, eps, scale=True, updates=None):
  """slim.batch_norm / tf.layers.batch_normalization(fused=True) over the last axis.

  Training: normalise with the biased batch variance; moving = moving*decay + batch*(1-decay)
  with the Bessel-corrected variance (TF fused batch norm).  `updates` collects
  the new moving statistics (the UPDATE_OPS the train op runs, abstract_model.py:335-381)."""
  beta = variables[scope + '/beta']
  gamma = variables[scope + '/gamma'] if scale else None
  red = tuple(range(x.dim() - 1))
  if training:
    mean = x.mean(red)
    var = x.var(red, unbiased=False)
    if updates is not None:
      n = x.numel() // x.shape[-1]
      mm = variables[scope + '/moving_mean']
      mv = variables[scope + '/moving_variance']
      updates[scope + '/moving_mean'] = (mm * decay + mean.detach() * (1 - decay))
      updates[scope + '/moving_variance'] = (mv * decay + var.detach() * (n / max(n - 1, 1)) * (1 - decay))
  else:
    mean = variables[scope + '/moving_mean']
    var = variables[scope + '/moving_variance']
  y = (x - mean) * torch.rsqrt(var + eps)
  if gamma is not None:
    y = y * gamma
  return _trace('bn', y + beta)


def dense(x, w_io, bias=None, fp32_path=False):
  """slim.fully_connected / tf.layers.dense: x @ W[in,out] + b.

  fp32_path marks the tiny layers the engine keeps in fp32 (action-context input FC, logits):
  under the bf16-storage model neither their weights nor their outputs are rounded."""
  if fp32_path:
    y = x @ w_io
    y = y if bias is None else y + bias
    if TRACE is not None:
      TRACE.append(('dense', y))
    return y
  y = x @ _store(w_io)
  return _trace('dense', y if bias is None else y + bias)


def log_loss(labels, predictions, eps=1e-7):
  """tf.losses.log_loss: mean over elements of -(y log(p+eps) + (1-y) log(1-p+eps))."""
  l = -(labels * torch.log(predictions + eps) + (1 - labels) * torch.log(1 - predictions + eps))
  return l.mean()


def l2_regularizer(scale, w):
  """slim.l2_regularizer: scale * sum(w^2) / 2 (tf.nn.l2_loss)."""
  return scale * (w * w).sum() / 2


def variance_scaling_std(fan_in):
  return math.sqrt(1.0 / fan_in) / .87962566103423978
