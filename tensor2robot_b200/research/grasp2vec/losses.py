"""Grasp2Vec embedding losses (research/grasp2vec/losses.py)."""
import torch

from tensor2robot_b200 import nn


class _AllGatherRows(torch.autograd.Function):
  """[B_local, D] -> [world * B_local, D] over the default process group; the backward hands every rank the rows of
  the gradient that belong to its own embeddings (each rank evaluates the same global loss)."""

  @staticmethod
  def forward(ctx, x):
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    ctx.rank, ctx.rows = rank, x.shape[0]
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous())
    return out

  @staticmethod
  def backward(ctx, grad):
    return grad[ctx.rank * ctx.rows:(ctx.rank + 1) * ctx.rows].contiguous()


def gather_global_batch(*embeddings):
  """SURVEY 8(e), opt-in: all-gathers [B_local, D] embeddings (2 MB per GPU at B = 256, D = 1024) so that a contrastive
  loss sees the GLOBAL batch as negatives.  Returns (gathered tensors, loss scale): every rank then evaluates the same
  global loss and back-propagates it through its own rows only, so the summed parameter gradient is the gradient of
  ONE global loss; the data-parallel step divides gradients by the world size, hence the loss is scaled by it.
  Without an initialised process group (or world size 1) this is the identity.  The reference never does this
  (single-replica loss, research/grasp2vec/grasp2vec_model.py:205-240)."""
  import torch.distributed as dist
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
    return embeddings, 1.0
  return tuple(_AllGatherRows.apply(e) for e in embeddings), float(dist.get_world_size())


def NPairsLoss(pregrasp_embedding, goal_embedding, postgrasp_embedding,  # pylint: disable=invalid-name
               non_negativity_constraint=False, global_negatives=False):
  """npairs_loss in both directions between (pre - post) and the goal embedding (losses.py:152-181).
  global_negatives=True: the negatives come from every replica's batch (gather_global_batch)."""
  pre, post, goal = (nn.to_f32(t) for t in (pregrasp_embedding, postgrasp_embedding, goal_embedding))
  pair_a = pre - post                 # [B, 1024] fp32: a host-scale elementwise op, left to torch autograd
  if non_negativity_constraint:
    pair_a = torch.relu(pair_a)
  pair_b = goal
  scale = 1.0
  if global_negatives:
    (pair_a, pair_b), scale = gather_global_batch(pair_a, pair_b)
  loss_1 = nn.npairs_loss(pair_a, pair_b)
  loss_2 = nn.npairs_loss(pair_b, pair_a)
  return (loss_1 + loss_2) * scale if scale != 1.0 else loss_1 + loss_2


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


# ---------------------------------------------------------------------------------------------
# Alternative / auxiliary embedding losses (losses.py:29-53, 80-157, 222-238).  [B, D] embedding-sized tails:
# plain torch on whatever device the embeddings live on.
# ---------------------------------------------------------------------------------------------
def _masked_mean(values, mask):
  """tf.dynamic_partition(values, mask, 2)[1] averaged; zeros(1) when the mask is empty (the tf.cond else-branch)."""
  mask = mask.reshape(-1).to(torch.int32)
  if int(mask.sum()) <= 0:
    return torch.zeros(1, dtype=torch.float32, device=values.device)
  return values[mask == 1].mean().float()


def L2ArithmeticLoss(pregrasp_embedding, goal_embedding, postgrasp_embedding, mask):  # pylint: disable=invalid-name
  """mean over the masked rows of ||pre - goal - post||^2 (losses.py:29-53)."""
  pre, post, goal = (nn.to_f32(t) for t in (pregrasp_embedding, postgrasp_embedding, goal_embedding))
  return _masked_mean(((pre - goal - post) ** 2).sum(1), mask)


def CosineArithmeticLoss(pregrasp_embedding, goal_embedding, postgrasp_embedding, mask):  # pylint: disable=invalid-name
  """mean over the masked rows of the cosine distance 1 - <l2n(pre - post), l2n(goal)> (losses.py:80-107)."""
  pre, post, goal = (nn.to_f32(t) for t in (pregrasp_embedding, postgrasp_embedding, goal_embedding))
  pair_a = torch.nn.functional.normalize(pre - post, dim=1, eps=1e-6)
  pair_b = torch.nn.functional.normalize(goal, dim=1, eps=1e-6)
  return _masked_mean(1.0 - (pair_a * pair_b).sum(1), mask)


def KeypointAccuracy(keypoints, labels):  # pylint: disable=invalid-name
  """Quadrant accuracy and sigmoid cross-entropy of spatial-softmax keypoints (Shapes dataset, losses.py:110-135)."""
  keypoints = nn.to_f32(keypoints).reshape(-1, 2)
  centers = torch.tensor([[0.5, -0.5], [-0.5, -0.5], [0.5, 0.5], [-0.5, 0.5]], dtype=torch.float32,
                         device=keypoints.device)
  logits = keypoints @ centers.t()
  labels = labels.reshape(-1).long()
  correct = (labels == torch.softmax(logits, 1).argmax(1)).float()
  onehot = torch.nn.functional.one_hot(labels, 4).float()
  loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, onehot)
  return correct.mean(), loss


def SendToZeroLoss(tensor, mask):  # pylint: disable=invalid-name
  """mean over the masked rows of ||tensor||_2 (losses.py:138-157)."""
  return _masked_mean(torch.linalg.norm(nn.to_f32(tensor), dim=1), mask)


def MatchNormsLoss(anchor_tensors, paired_tensors):  # pylint: disable=invalid-name
  """tf.nn.l2_loss of the row-norm differences (sum d^2 / 2); gradients reach only the paired tensors
  (losses.py:222-238)."""
  anchor_norms = torch.linalg.norm(nn.to_f32(anchor_tensors), dim=1).detach()
  paired_norms = torch.linalg.norm(nn.to_f32(paired_tensors), dim=1)
  return ((anchor_norms - paired_norms) ** 2).sum() / 2


def NPairsLossMultilabel(pregrasp_embedding, goal_embedding, postgrasp_embedding, grasp_success, params=None):  # pylint: disable=invalid-name
  """npairs_loss_multilabel in both directions (losses.py:188-219).  Example i carries the single class i * success_i of
  B + 1 classes, so every failed grasp shares class 0 (with example 0 as well); slim's loss then is the softmax
  cross-entropy of the similarity matrix against the row-normalised label adjacency (shared classes between examples
  i and j) plus the embedding regulariser 0.25 * 0.002 * (mean |a|^2 + mean |p|^2).  With every grasp successful the
  adjacency is the identity and this equals NPairsLoss (the reference's own test)."""
  del params
  pre, post, goal = (nn.to_f32(t) for t in (pregrasp_embedding, postgrasp_embedding, goal_embedding))
  pair_a, pair_b = pre - post, goal
  b = pre.shape[0]
  success = torch.as_tensor(grasp_success, device=pre.device).reshape(-1).long()
  classes = torch.arange(b, device=pre.device) * success
  labels = torch.nn.functional.one_hot(classes, b + 1).float()
  adjacency = labels @ labels.t()
  target = adjacency / adjacency.sum(1, keepdim=True)

  def one_direction(anchor, positive):
    reg = 0.25 * 0.002 * ((anchor ** 2).sum(1).mean() + (positive ** 2).sum(1).mean())
    similarity = anchor @ positive.t()
    return reg - (target * torch.log_softmax(similarity, dim=1)).sum(1).mean()

  return one_direction(pair_a, pair_b) + one_direction(pair_b, pair_a)


def _GetSoftMaxResponse(goal_embedding, scene_spatial):  # pylint: disable=invalid-name
  """Heat map <scene_spatial[b, y, x, :], goal_embedding[b]>: its maximum and the maximum of its spatial softmax
  (losses.py:241-266)."""
  goal, scene = nn.to_f32(goal_embedding), nn.to_f32(scene_spatial)
  b = goal.shape[0]
  heat = (scene * goal.reshape(b, 1, 1, -1)).sum(3).reshape(b, -1)
  return heat.max(1).values, torch.softmax(heat, dim=1).max(1).values


def TYloss(pregrasp_spatial, postgrasp_spatial, goal_embedding):  # pylint: disable=invalid-name
  """mean over the batch of (max cosine response of the goal in the postgrasp map) - (the same in the pregrasp map):
  the object should be found before the grasp and not after it (losses.py:269-303)."""
  l2n = lambda t: torch.nn.functional.normalize(nn.to_f32(t), dim=-1, eps=1e-6)
  pre, post, goal = l2n(pregrasp_spatial), l2n(postgrasp_spatial), l2n(goal_embedding)
  goal = goal[:, None, None, :]
  pre_max = (pre * goal).sum(-1).flatten(1).max(1).values
  post_max = (post * goal).sum(-1).flatten(1).max(1).values
  return (post_max - pre_max).mean()
