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


def _np_npairs(anchor, positive, target):
  reg = 0.25 * 0.002 * ((anchor ** 2).sum(1).mean() + (positive ** 2).sum(1).mean())
  sim = anchor @ positive.T
  logp = sim - sim.max(1, keepdims=True)
  logp = logp - np.log(np.exp(logp).sum(1, keepdims=True))
  return reg - (target * logp).sum(1).mean()


def test_reference_npairs_multilabel_case():
  """research/grasp2vec/losses_test.py:150-177: with every grasp successful the multilabel loss equals the n-pairs loss
  (to 5 places); with failures it is larger than both.  Values against the numpy restatement of slim's
  npairs_loss_multilabel (softmax cross-entropy against the row-normalised label adjacency + embedding regulariser)."""
  rng = np.random.RandomState(0)
  a, c, small = rng.rand(16), rng.rand(16), rng.rand(16) / 10
  pre = np.array([a + small, a, c]).astype(np.float32)
  post = np.zeros_like(pre)
  goal = pre.copy()
  t = torch.from_numpy

  def both(target):
    pa = (pre - post).astype(np.float64)
    return _np_npairs(pa, goal.astype(np.float64), target) + _np_npairs(goal.astype(np.float64), pa, target)

  all_ok = float(losses.NPairsLossMultilabel(t(pre), t(goal), t(post), np.ones(3, np.int32), {}))
  single = both(np.eye(3))
  assert abs(all_ok - single) < 1e-5                                   # == NPairsLoss (identity adjacency)
  success = np.array([0, 0, 1])
  failed = float(losses.NPairsLossMultilabel(t(pre), t(goal), t(post), success, {}))
  labels = np.eye(4)[np.arange(3) * success]
  adjacency = labels @ labels.T
  assert abs(failed - both(adjacency / adjacency.sum(1, keepdims=True))) < 1e-5
  assert failed > single and failed > all_ok


def test_soft_max_response_and_ty_loss():
  """losses.py:241-303 against direct numpy restatements; TYloss is negative when the goal is present before the grasp
  and absent after it."""
  rng = np.random.RandomState(1)
  b, h, w, d = 3, 4, 5, 6
  goal = rng.standard_normal((b, d)).astype(np.float32)
  pre = rng.standard_normal((b, h, w, d)).astype(np.float32)
  post = rng.standard_normal((b, h, w, d)).astype(np.float32)
  t = torch.from_numpy
  max_heat, max_soft = losses._GetSoftMaxResponse(t(goal), t(pre))    # pylint: disable=protected-access
  heat = (pre * goal[:, None, None, :]).sum(3).reshape(b, -1)
  soft = np.exp(heat - heat.max(1, keepdims=True))
  soft /= soft.sum(1, keepdims=True)
  np.testing.assert_allclose(max_heat.numpy(), heat.max(1), rtol=1e-5)
  np.testing.assert_allclose(max_soft.numpy(), soft.max(1), rtol=1e-5)
  n = lambda x: x / np.sqrt(np.maximum((x ** 2).sum(-1, keepdims=True), 1e-12))
  want = ((n(post) * n(goal)[:, None, None, :]).sum(-1).reshape(b, -1).max(1) -
          (n(pre) * n(goal)[:, None, None, :]).sum(-1).reshape(b, -1).max(1)).mean()
  np.testing.assert_allclose(float(losses.TYloss(t(pre), t(post), t(goal))), want, rtol=1e-5, atol=1e-6)
  pre[:, 1, 2, :] = goal * 3.0                                         # the object is in the pregrasp scene ...
  post = -np.abs(post) * np.sign(goal)[:, None, None, :]               # ... and every postgrasp location points away
  assert float(losses.TYloss(t(pre), t(post), t(goal))) < -1.0
