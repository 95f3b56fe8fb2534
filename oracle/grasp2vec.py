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
