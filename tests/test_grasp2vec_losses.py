"""Grasp2Vec auxiliary embedding losses (research/grasp2vec/losses.py:29-157, 222-238) against numpy restatements of
the TF expressions; [B, D] tails that run on any torch device."""
import numpy as np
import torch

from tensor2robot_b200.research.grasp2vec import losses


def _l2n(x):
  return x / np.sqrt(np.maximum((x ** 2).sum(1, keepdims=True), 1e-12))


def test_arithmetic_losses():
  rng = np.random.RandomState(0)
  pre, goal, post = (rng.standard_normal((6, 16)).astype(np.float32) for _ in range(3))
  mask = np.array([1, 0, 1, 1, 0, 0])
  t = lambda a: torch.from_numpy(a)
  want = ((pre - goal - post) ** 2).sum(1)[mask == 1].mean()
  np.testing.assert_allclose(float(losses.L2ArithmeticLoss(t(pre), t(goal), t(post), t(mask))), want, rtol=1e-6)
  cos = 1 - (_l2n(pre - post) * _l2n(goal)).sum(1)
  np.testing.assert_allclose(float(losses.CosineArithmeticLoss(t(pre), t(goal), t(post), t(mask.reshape(6, 1)))),
                             cos[mask == 1].mean(), rtol=1e-5)
  np.testing.assert_allclose(float(losses.SendToZeroLoss(t(pre), t(mask))), np.linalg.norm(pre, axis=1)[mask == 1].mean(),
                             rtol=1e-6)
  empty = torch.zeros(6, dtype=torch.int64)
  for out in (losses.L2ArithmeticLoss(t(pre), t(goal), t(post), empty), losses.SendToZeroLoss(t(pre), empty),
              losses.CosineArithmeticLoss(t(pre), t(goal), t(post), empty)):
    assert out.shape == (1,) and float(out) == 0.0          # tf.zeros(1) in the reference's else-branch


def test_match_norms_and_keypoints():
  rng = np.random.RandomState(1)
  a = torch.from_numpy(rng.standard_normal((5, 8)).astype(np.float32)).requires_grad_(True)
  p = torch.from_numpy(rng.standard_normal((5, 8)).astype(np.float32)).requires_grad_(True)
  loss = losses.MatchNormsLoss(a, p)
  d = np.linalg.norm(a.detach().numpy(), axis=1) - np.linalg.norm(p.detach().numpy(), axis=1)
  np.testing.assert_allclose(float(loss), (d ** 2).sum() / 2, rtol=1e-6)
  loss.backward()
  assert a.grad is None and p.grad.abs().max() > 0           # stop_gradient on the anchors
  kp = np.array([[0.6, -0.4], [-0.7, -0.2], [0.3, 0.9], [-0.5, 0.5], [0.4, 0.4]], np.float32)
  labels = np.array([0, 1, 2, 3, 0])
  acc, ce = losses.KeypointAccuracy(torch.from_numpy(kp.reshape(5, 1, 2)), torch.from_numpy(labels))
  centers = np.array([[0.5, -0.5], [-0.5, -0.5], [0.5, 0.5], [-0.5, 0.5]], np.float32)
  logits = kp @ centers.T
  onehot = np.eye(4, dtype=np.float32)[labels]
  want_ce = (np.maximum(logits, 0) - logits * onehot + np.log1p(np.exp(-np.abs(logits)))).mean()
  np.testing.assert_allclose(float(acc), 0.8, rtol=1e-6)      # the last keypoint lies in quadrant 2, labelled 0
  np.testing.assert_allclose(float(ce), want_ce, rtol=1e-6)


# the reference's own cases (research/grasp2vec/losses_test.py:46-148)
import pytest  # noqa: E402

_RNG = np.random.RandomState(7)
_FAKE = {k: _RNG.random_sample((32, 8)).astype(np.float32) for k in ('pre', 'post', 'goal')}


def _cos(x, y):
  return 1 - (x * y).sum(1) / (np.linalg.norm(x, axis=1) * np.linalg.norm(y, axis=1))


@pytest.mark.parametrize('mask_kind', ['zeros', 'ones', 'mixed'])
def test_reference_arithmetic_loss_cases(mask_kind):
  mask = {'zeros': np.zeros(32), 'ones': np.ones(32), 'mixed': np.eye(1, 32)[0]}[mask_kind]
  t = torch.from_numpy
  cosine = float(losses.CosineArithmeticLoss(t(_FAKE['pre']), t(_FAKE['goal']), t(_FAKE['post']), t(mask)))
  l2 = float(losses.L2ArithmeticLoss(t(_FAKE['pre']), t(_FAKE['goal']), t(_FAKE['post']), t(mask)))
  if mask_kind == 'zeros':
    assert cosine == 0 and l2 == 0
    return
  rows = slice(None) if mask_kind == 'ones' else slice(0, 1)
  want_cos = _cos(_FAKE['pre'][rows] - _FAKE['post'][rows], _FAKE['goal'][rows]).mean()
  want_l2 = (np.linalg.norm(_FAKE['pre'][rows] - (_FAKE['post'][rows] + _FAKE['goal'][rows]), axis=1) ** 2).mean()
  assert abs(cosine - want_cos) < 1e-3 and abs(l2 - want_l2) < 1e-3           # assertAlmostEqual(places=3)


@pytest.mark.parametrize('keypoints,labels,expected_accuracy', [
    ([[-0.5, -0.5], [-0.5, 0.5], [0.5, -0.5], [0.5, 0.5]], [1, 3, 0, 2], 1.0),          # CorrectKeypoints
    ([[-0.5, -0.5], [-0.5, 0.5], [0.5, -0.5], [0.5, 0.5]], [2, 0, 3, 1], 0.0),          # IncorrectKeypoints
    ([[-0.6, -0.4], [-0.4, 0.1], [0.3, -0.6], [0.7, 0.9]], [1, 0, 0, 1], 0.5),          # HalfcorrectKeypoints
])
def test_reference_keypoint_accuracy_cases(keypoints, labels, expected_accuracy):
  acc, _ = losses.KeypointAccuracy(torch.tensor(keypoints, dtype=torch.float32), torch.tensor(labels))
  assert float(acc) == expected_accuracy
