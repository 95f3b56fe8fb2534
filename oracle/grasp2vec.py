"""TEST INFRASTRUCTURE ONLY (oracle): Grasp2Vec's n-pairs loss in torch (differentiable on CPU).

Restates research/grasp2vec/losses.py:152-181 (NPairsLoss).  The inner loss is a third-party dependency that
is absent from /root/reference (tf.contrib.losses.metric_learning.npairs_loss, TensorFlow 1.15 contrib /
tf_slim `metric_learning.npairs_loss`); its published algorithm:
    reg  = 0.25 * reg_lambda * (mean_i |a_i|^2 + mean_i |p_i|^2),   reg_lambda = 0.002
    sim  = a @ p^T;  labels_remapped = (labels == labels^T) / row sum  (identity for labels = range(B))
    xent = mean_i softmax_cross_entropy(sim_i, labels_remapped_i)
Parity unpinned against TensorFlow itself (absent); anchored on the formula above and on
research/grasp2vec/losses_test.py's use (finite scalar loss for random 32 x 512 embeddings)."""
import torch


def npairs_loss(anchor, positive, reg_lambda=0.002):
  reg_anchor = (anchor ** 2).sum(1).mean()
  reg_positive = (positive ** 2).sum(1).mean()
  l2loss = 0.25 * reg_lambda * (reg_anchor + reg_positive)
  sim = anchor @ positive.t()
  labels = torch.eye(anchor.shape[0], dtype=sim.dtype)
  xent = -(labels * torch.log_softmax(sim, dim=1)).sum(1).mean()
  return l2loss + xent


def npairs_loss_both(pregrasp, goal, postgrasp, non_negativity_constraint=False):
  pair_a = pregrasp - postgrasp
  if non_negativity_constraint:
    pair_a = torch.relu(pair_a)
  return npairs_loss(pair_a, goal) + npairs_loss(goal, pair_a)


def pairwise_distance_squared(e):
  """tf.contrib metric_learning.pairwise_distance(squared=True): |a|^2 + |b|^2 - 2ab clamped at 0, zero diagonal."""
  sq = (e ** 2).sum(1, keepdim=True)
  d = torch.clamp(sq + sq.t() - 2.0 * e @ e.t(), min=0.0)
  return d * (1.0 - torch.eye(e.shape[0], dtype=e.dtype))


def triplet_semihard_loss(labels, embeddings, margin=1.0):
  """The mining of layers/tec.py:322-383 (a reproduction of tf-slim's triplet_semihard_loss) with the squared
  Euclidean distance tf-slim uses; written pair by pair, differentiable through torch."""
  d = pairwise_distance_squared(embeddings)
  m = d.shape[0]
  total, num = 0.0, 0
  for a in range(m):
    neg = [n for n in range(m) if labels[n] != labels[a]]
    for p in range(m):
      if p == a or labels[p] != labels[a]:
        continue
      num += 1
      if not neg:
        continue
      dn = d[a, neg]
      outside = dn[dn > d[a, p]]
      semi_hard = outside.min() if outside.numel() else dn.max()
      total = total + torch.clamp(margin + d[a, p] - semi_hard, min=0.0)
  return total / max(num, 1)


def triplet_loss(pregrasp, goal, postgrasp):
  pair_a = torch.nn.functional.normalize(pregrasp - postgrasp, dim=1, eps=1e-6)
  pair_b = torch.nn.functional.normalize(goal, dim=1, eps=1e-6)
  labels = list(range(pregrasp.shape[0])) * 2
  return triplet_semihard_loss(labels, torch.cat([pair_a, pair_b], 0), margin=3.0)
