"""TEST INFRASTRUCTURE ONLY - numpy float64 restatement of MockT2RModel.inference_network_fn (utils/mocks.py:150-176):
3 x (tf.layers.dense + elu, tf.layers.batch_normalization with training=False, epsilon 1e-3) -> dense(1).
Pinned with real TensorFlow weights: the reference's fixture checkpoint (trained 1100 steps on MockInputGenerator's
linearly separable data) must separate that dataset through this restatement (tests/test_mocks.py)."""
import numpy as np


def forward(weights, x, eps=1e-3):
  net = np.asarray(x, np.float64)
  for pos in range(3):
    d, b = 'MockT2RModel.dense.%d' % pos, 'MockT2RModel.batch_norm.%d' % pos
    net = net @ np.asarray(weights[d + '/kernel'], np.float64) + np.asarray(weights[d + '/bias'], np.float64)
    net = np.where(net > 0, net, np.expm1(np.minimum(net, 0)))
    net = ((net - weights[b + '/moving_mean']) / np.sqrt(np.asarray(weights[b + '/moving_variance'], np.float64) + eps)
           * weights[b + '/gamma'] + weights[b + '/beta'])
  return net @ np.asarray(weights['MockT2RModel.dense.4/kernel'], np.float64) + weights['MockT2RModel.dense.4/bias']


def categorical_hinge(y_true, y_pred):
  pos = (y_true * y_pred).sum(-1)
  neg = ((1.0 - y_true) * y_pred).max(-1)
  return np.maximum(neg - pos + 1.0, 0.0).mean()
