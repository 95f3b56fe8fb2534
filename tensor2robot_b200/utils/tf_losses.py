"""tf.losses restated on torch tensors (the label-sized [B, k] tail of a model: host-scale logic next to the
networks).  Reduction.SUM_BY_NONZERO_WEIGHTS everywhere, like the TF1 defaults."""
import torch


def compute_weighted_loss(loss, weights=1.0):
  """tf.losses.compute_weighted_loss: sum(loss * w) / #{w != 0} with w broadcast to the loss."""
  w = torch.as_tensor(weights, dtype=loss.dtype, device=loss.device)
  w = torch.broadcast_to(w, loss.shape)
  nonzero = (w != 0).sum().to(loss.dtype)
  # no non-zero weight => the numerator is zero too: the safe division needs no host-side branch (and no sync)
  return (loss * w).sum() / torch.clamp(nonzero, min=1.0)


def huber_loss(labels, predictions, weights=1.0, delta=1.0):
  """tf.losses.huber_loss."""
  err = (predictions - labels).abs()
  quad = torch.clamp(err, max=delta)
  return compute_weighted_loss(0.5 * quad ** 2 + delta * (err - quad), weights)


def mean_squared_error(labels, predictions, weights=1.0):
  """tf.losses.mean_squared_error."""
  return compute_weighted_loss((predictions - labels) ** 2, weights)


def log_loss(labels, predictions, weights=1.0, epsilon=1e-7):
  """tf.losses.log_loss."""
  return compute_weighted_loss(
      -labels * torch.log(predictions + epsilon) - (1 - labels) * torch.log(1 - predictions + epsilon), weights)
