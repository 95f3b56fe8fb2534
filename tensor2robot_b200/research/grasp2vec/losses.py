"""Grasp2Vec embedding losses (research/grasp2vec/losses.py)."""
import torch

from tensor2robot_b200 import nn


def NPairsLoss(pregrasp_embedding, goal_embedding, postgrasp_embedding,  # pylint: disable=invalid-name
               non_negativity_constraint=False):
  """npairs_loss in both directions between (pre - post) and the goal embedding (losses.py:152-181)."""
  pre, post, goal = (nn.to_f32(t) for t in (pregrasp_embedding, postgrasp_embedding, goal_embedding))
  pair_a = pre - post                 # [B, 1024] fp32: a host-scale elementwise op, left to torch autograd
  if non_negativity_constraint:
    pair_a = torch.relu(pair_a)
  pair_b = goal
  loss_1 = nn.npairs_loss(pair_a, pair_b)
  loss_2 = nn.npairs_loss(pair_b, pair_a)
  return loss_1 + loss_2


def TripletLoss(pregrasp_embedding, goal_embedding, postgrasp_embedding):  # pylint: disable=invalid-name
  """Semi-hard mining triplet loss between l2-normalised (pre - post) and goal embeddings, labels
  tile(range(B), 2), margin 3.0 (losses.py:51-71)."""
  pre, post, goal = (nn.to_f32(t) for t in (pregrasp_embedding, postgrasp_embedding, goal_embedding))
  pair_a = torch.nn.functional.normalize(pre - post, dim=1, eps=1e-6)    # tf.nn.l2_normalize: x * rsqrt(max(sum x^2, 1e-12))
  pair_b = torch.nn.functional.normalize(goal, dim=1, eps=1e-6)
  b = pre.shape[0]
  labels = torch.arange(b, dtype=torch.int32).repeat(2)
  pairs = torch.cat([pair_a, pair_b], dim=0)
  return nn.triplet_semihard_loss(labels, pairs, margin=3.0)
