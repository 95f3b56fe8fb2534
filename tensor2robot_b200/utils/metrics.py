"""Streaming evaluation metrics with the semantics of tf.metrics.* (what the reference's `model_eval_fn`s return,
models/abstract_model.py:506-565, research/bcz/model.py:588-617, 894-929): every call produces the sufficient
statistics of ONE batch as device tensors; the evaluation loop adds the statistics of all batches and finishes the value
once (a ratio of sums, not a mean of per-batch ratios).  Nothing here synchronises the host."""
import torch


class Metric(object):
  """stats: tuple of tensors that add across batches; finish(*stats) -> scalar tensor."""

  def __init__(self, stats, finish):
    self.stats = tuple(stats)
    self.finish = finish

  def merge(self, other):
    return Metric(tuple(a + b for a, b in zip(self.stats, other.stats)), self.finish)

  def result(self):
    return self.finish(*self.stats)


def _f(x):
  return x if torch.is_tensor(x) else torch.as_tensor(x)


def _safe_div(num, den):
  """tf's div_no_nan / `safe_div`: 0 where the denominator is 0."""
  return torch.where(den > 0, num / torch.clamp(den, min=1e-38), torch.zeros_like(num))


def mean(values, weights=None):
  """tf.metrics.mean: sum(values * weights) / sum(weights) over every element seen."""
  v = _f(values).detach().double()
  if weights is None:
    total, count = v.sum(), torch.as_tensor(float(v.numel()), dtype=torch.float64, device=v.device)
  else:
    w = torch.broadcast_to(_f(weights).detach().double().to(v.device), v.shape)
    total, count = (v * w).sum(), w.sum()
  return Metric((total, count), _safe_div)


def accuracy(labels, predictions, weights=None):
  """tf.metrics.accuracy: the frequency of predictions == labels."""
  l, p = _f(labels).detach(), _f(predictions).detach()
  if p.dtype != l.dtype:
    p = p.to(l.dtype)
  return mean((l == p).double(), weights)


def _confusion(labels, predictions):
  l, p = _f(labels).detach() != 0, _f(predictions).detach() != 0
  s = lambda m: m.double().sum()
  return s(l & p), s(~l & p), s(l & ~p)      # true positives, false positives, false negatives


def precision(labels, predictions):
  """tf.metrics.precision on values cast to bool: tp / (tp + fp), 0 when nothing was predicted positive."""
  tp, fp, _ = _confusion(labels, predictions)
  return Metric((tp, fp), lambda a, b: _safe_div(a, a + b))


def recall(labels, predictions):
  """tf.metrics.recall: tp / (tp + fn)."""
  tp, _, fn = _confusion(labels, predictions)
  return Metric((tp, fn), lambda a, b: _safe_div(a, a + b))


def auc(labels, predictions, num_thresholds=200):
  """tf.metrics.auc(curve='ROC', summation_method='trapezoidal'): confusion counts at `num_thresholds` thresholds
  ({-1e-7, 1/(n-1), ..., (n-2)/(n-1), 1 + 1e-7}; a prediction is positive when it is > the threshold), accumulated over
  the batches; area under (false positive rate, recall) by the trapezoid rule with tf's epsilon 1e-6."""
  l = (_f(labels).detach() != 0).reshape(-1)
  p = _f(predictions).detach().double().reshape(-1)
  n = num_thresholds
  inner = torch.arange(1, n - 1, dtype=torch.float64, device=p.device) / (n - 1)
  thresholds = torch.cat([torch.full((1,), -1e-7, dtype=torch.float64, device=p.device), inner,
                          torch.full((1,), 1.0 + 1e-7, dtype=torch.float64, device=p.device)])
  positive = p[None, :] > thresholds[:, None]                          # [thresholds, elements]
  lab = l[None, :]
  s = lambda m: m.double().sum(1)
  stats = (s(positive & lab), s(positive & ~lab), s(~positive & lab), s(~positive & ~lab))   # tp, fp, fn, tn

  def finish(tp, fp, fn, tn):
    eps = 1e-6
    rec = (tp + eps) / (tp + fn + eps)
    fpr = fp / (fp + tn + eps)
    return ((fpr[:-1] - fpr[1:]) * (rec[:-1] + rec[1:]) / 2.0).sum()

  return Metric(stats, finish)


class Accumulator(object):
  """What the evaluation loop keeps: {name: Metric} merged batch by batch; plain tensors / numbers are averaged over the
  batches (the default `{'loss': train_loss}` of AbstractT2RModel.model_eval_fn)."""

  def __init__(self):
    self._metrics = {}

  def update(self, metrics):
    for name, value in (metrics or {}).items():
      if value is None:
        continue
      if not isinstance(value, Metric):
        value = mean(_f(value).detach().double().mean())
      self._metrics[name] = value if name not in self._metrics else self._metrics[name].merge(value)

  def results(self):
    return {name: float(m.result()) for name, m in self._metrics.items()}
