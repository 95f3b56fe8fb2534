"""utils/metrics.py: the streaming semantics of tf.metrics.* (statistics summed over batches, ratio taken once) and the
ROC-AUC formula of tf.metrics.auc (200 thresholds, trapezoid, epsilon 1e-6) against direct numpy restatements."""
import numpy as np
import torch

from tensor2robot_b200.utils import metrics


def _stream(fn, batches):
  acc = metrics.Accumulator()
  for args in batches:
    acc.update({'m': fn(*[torch.from_numpy(np.asarray(a)) for a in args])})
  return acc.results()['m']


def test_mean_accuracy_precision_recall_are_ratios_of_sums():
  rng = np.random.RandomState(0)
  batches = [(rng.randint(0, 2, n).astype(np.float32), rng.randint(0, 2, n).astype(np.float32)) for n in (3, 17, 64)]
  labels = np.concatenate([b[0] for b in batches])
  preds = np.concatenate([b[1] for b in batches])
  assert abs(_stream(metrics.accuracy, batches) - (labels == preds).mean()) < 1e-12
  tp = ((labels == 1) & (preds == 1)).sum()
  assert abs(_stream(metrics.precision, batches) - tp / (preds == 1).sum()) < 1e-12
  assert abs(_stream(metrics.recall, batches) - tp / (labels == 1).sum()) < 1e-12
  values = [(rng.standard_normal(n).astype(np.float32),) for n in (5, 1, 30)]      # unequal batches: not a mean of means
  assert abs(_stream(metrics.mean, values) - np.concatenate([v[0] for v in values]).mean()) < 1e-6
  # nothing predicted positive / no positive label: 0, not NaN (tf's div_no_nan)
  zeros = [(np.ones(4, np.float32), np.zeros(4, np.float32))]
  assert _stream(metrics.precision, zeros) == 0.0
  assert _stream(metrics.recall, [(np.zeros(4, np.float32), np.ones(4, np.float32))]) == 0.0


def test_auc_matches_the_tf_formula():
  rng = np.random.RandomState(1)
  batches = [(rng.randint(0, 2, n).astype(np.float32), rng.uniform(0, 1, n).astype(np.float32)) for n in (50, 7, 93)]
  labels = np.concatenate([b[0] for b in batches]).astype(bool)
  preds = np.concatenate([b[1] for b in batches]).astype(np.float64)
  n = 200
  thresholds = np.array([-1e-7] + [(i + 1) / (n - 1) for i in range(n - 2)] + [1 + 1e-7])
  tp = np.array([((preds > t) & labels).sum() for t in thresholds], np.float64)
  fp = np.array([((preds > t) & ~labels).sum() for t in thresholds], np.float64)
  fn = np.array([((preds <= t) & labels).sum() for t in thresholds], np.float64)
  tn = np.array([((preds <= t) & ~labels).sum() for t in thresholds], np.float64)
  rec, fpr = (tp + 1e-6) / (tp + fn + 1e-6), fp / (fp + tn + 1e-6)
  want = ((fpr[:-1] - fpr[1:]) * (rec[:-1] + rec[1:]) / 2).sum()
  got = _stream(metrics.auc, batches)
  assert abs(got - want) < 1e-9
  # sanity: close to the exact rank statistic, 1 for a perfect and 0.5 for a constant predictor
  pos, neg = preds[labels], preds[~labels]
  exact = (pos[:, None] > neg[None, :]).mean()
  assert abs(got - exact) < 2e-2
  assert abs(_stream(metrics.auc, [(labels.astype(np.float32), labels.astype(np.float32))]) - 1.0) < 1e-4
  assert abs(_stream(metrics.auc, [(labels.astype(np.float32), np.full(labels.shape, 0.3, np.float32))]) - 0.5) < 1e-4


def test_accumulator_averages_plain_values_over_batches():
  acc = metrics.Accumulator()
  for v in (1.0, 2.0, 6.0):
    acc.update({'loss': torch.tensor(v), 'skipped': None})
  assert acc.results() == {'loss': 3.0}
