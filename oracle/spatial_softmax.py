"""TEST INFRASTRUCTURE ONLY (oracle): numpy restatement of layers/spatial_softmax.py:29-88
(BuildSpatialSoftmax, deterministic branch).  Parity unpinned: the reference needs TensorFlow, which is
absent; the statement-by-statement restatement below is the anchor (note the interleaved output)."""
import numpy as np


def build_spatial_softmax(features):
  """features [B, H, W, C] float -> (expected_feature_points [B, 2C], softmax [B, H, W, C])."""
  features = np.asarray(features, np.float64)
  b, num_rows, num_cols, num_features = features.shape
  x_pos = np.empty([num_rows, num_cols], np.float32)                        # :51-58
  y_pos = np.empty([num_rows, num_cols], np.float32)
  for i in range(num_rows):
    for j in range(num_cols):
      x_pos[i, j] = 2.0 * j / (num_cols - 1.0) - 1.0
      y_pos[i, j] = 2.0 * i / (num_rows - 1.0) - 1.0
  x_pos = x_pos.reshape(num_rows * num_cols)                                # :60-61
  y_pos = y_pos.reshape(num_rows * num_cols)
  feats = features.transpose(0, 3, 1, 2).reshape(-1, num_rows * num_cols)   # :67-68
  e = np.exp(feats - feats.max(axis=1, keepdims=True))                      # tf.nn.softmax :77
  softmax = e / e.sum(axis=1, keepdims=True)
  x_output = (x_pos * softmax).sum(axis=1, keepdims=True)                   # :79-83
  y_output = (y_pos * softmax).sum(axis=1, keepdims=True)
  points = np.concatenate([x_output, y_output], 1).reshape(-1, num_features * 2)   # :85-86 (interleaved!)
  heat = softmax.reshape(-1, num_features, num_rows, num_cols).transpose(0, 2, 3, 1)  # :87-89
  return points.astype(np.float32), heat.astype(np.float32)
