"""Grasp2Vec (SURVEY A-27): n-pairs loss kernel against the oracle formula (value + gradients), and one
training step of the model (two truncated ResNet-50 towers) end to end."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('b,d', [(32, 512), (7, 64), (256, 1024)])
def test_npairs_loss_matches_oracle(b, d):
  from oracle import grasp2vec as oracle
  from tensor2robot_b200 import nn
  rng = np.random.RandomState(b)
  a = torch.from_numpy(rng.uniform(0, 1, (b, d)).astype(np.float32) * 0.3)
  p = torch.from_numpy(rng.uniform(0, 1, (b, d)).astype(np.float32) * 0.3)
  ao, po = a.clone().requires_grad_(True), p.clone().requires_grad_(True)
  lo = oracle.npairs_loss(ao.double(), po.double())
  lo.backward()
  ag, pg = a.cuda().requires_grad_(True), p.cuda().requires_grad_(True)
  l = nn.npairs_loss(ag, pg)
  (l * 1.5).backward()
  assert abs(l.item() - lo.item()) < 2e-5 * max(1.0, abs(lo.item()))
  np.testing.assert_allclose(ag.grad.cpu().numpy(), 1.5 * ao.grad.numpy(), rtol=2e-4, atol=2e-6)
  np.testing.assert_allclose(pg.grad.cpu().numpy(), 1.5 * po.grad.numpy(), rtol=2e-4, atol=2e-6)


def test_grasp2vec_losses_test_semantics():
  """research/grasp2vec/losses_test.py: NPairsLoss on random (32, 512) embeddings is a finite scalar; both
  directions are summed; the non-negativity variant applies ReLU to pre - post."""
  from oracle import grasp2vec as oracle
  from tensor2robot_b200.research.grasp2vec import losses
  rng = np.random.RandomState(0)
  pre, goal, post = (torch.from_numpy(rng.uniform(size=(32, 512)).astype(np.float32)) for _ in range(3))
  for nonneg in (False, True):
    got = losses.NPairsLoss(pre.cuda(), goal.cuda(), post.cuda(), non_negativity_constraint=nonneg)
    want = oracle.npairs_loss_both(pre.double(), goal.double(), post.double(), nonneg)
    assert got.dim() == 0 and np.isfinite(got.item())
    assert abs(got.item() - want.item()) < 1e-4 * abs(want.item())


def test_grasp2vec_model_train_step():
  from tensor2robot_b200 import nn
  from tensor2robot_b200.research.grasp2vec import grasp2vec_model
  from tensor2robot_b200.utils import tensorspec_utils
  model = grasp2vec_model.Grasp2VecModel(scene_size=(472, 472), goal_size=(472, 472))
  pre = model.preprocessor
  in_spec = pre.get_in_feature_specification('train')
  assert tuple(in_spec['pregrasp_image'].shape) == (512, 640, 3) and in_spec['goal_image'].data_format == 'jpeg'
  rng = np.random.RandomState(1)
  b = 4
  feats = tensorspec_utils.TensorSpecStruct()
  for name in ('pregrasp_image', 'postgrasp_image', 'goal_image'):
    feats[name] = torch.from_numpy(rng.randint(0, 256, (b, 512, 640, 3)).astype(np.uint8)).cuda()
  grasp2vec_model.seed(3)
  feats, _ = pre._preprocess_fn(feats, None, 'train')
  assert tuple(feats.pregrasp_image.shape) == (b, 472, 472, 3) and feats.goal_image.dtype == torch.bfloat16
  vs = nn.VariableStore('cuda')
  with torch.no_grad(), nn.variable_store(vs):
    model.inference_network_fn(feats, None, 'eval')
  vs.finalize()
  names = list(vs.export_tf().keys())
  assert any(n.startswith('scene/resnet_model/') for n in names) and any(n.startswith('goal/resnet_model/') for n in names)
  with nn.variable_store(vs):
    out = model.inference_network_fn(feats, None, 'train')
    assert tuple(out['pre_vector'].shape) == (b, 1024) and tuple(out['goal_spatial'].shape)[0] == b
    loss, train_outputs = model.model_train_fn(feats, None, out, 'train')
    vs.zero_grad()
    loss.backward()
  torch.cuda.synchronize()
  assert np.isfinite(loss.item()) and 'embed_loss' in train_outputs
  g = vs.flat_grad
  assert torch.isfinite(g).all() and float(g.abs().max()) > 0


@pytest.mark.parametrize('m,d,nlabels,margin', [(16, 32, 8, 1.0), (64, 128, 32, 3.0), (48, 64, 5, 0.2)])
def test_triplet_semihard_loss_matches_oracle(m, d, nlabels, margin):
  from oracle import grasp2vec as oracle
  from tensor2robot_b200 import nn
  rng = np.random.RandomState(m + d)
  e = torch.from_numpy(rng.standard_normal((m, d)).astype(np.float32) * 0.4)
  labels = [int(v) for v in rng.randint(0, nlabels, m)]
  eo = e.clone().double().requires_grad_(True)
  lo = oracle.triplet_semihard_loss(labels, eo, margin)
  lo.backward()
  eg = e.cuda().requires_grad_(True)
  l = nn.triplet_semihard_loss(torch.tensor(labels), eg, margin)
  l.backward()
  assert abs(l.detach().item() - float(lo)) < 1e-4 * max(1.0, abs(float(lo))), (l.detach().item(), float(lo))
  np.testing.assert_allclose(eg.grad.cpu().numpy(), eo.grad.numpy(), rtol=2e-3, atol=2e-5)


def test_grasp2vec_triplet_loss():
  from oracle import grasp2vec as oracle
  from tensor2robot_b200.research.grasp2vec import losses
  rng = np.random.RandomState(2)
  pre, goal, post = (torch.from_numpy(rng.uniform(size=(32, 512)).astype(np.float32)) for _ in range(3))
  got = losses.TripletLoss(pre.cuda(), goal.cuda(), post.cuda())
  want = oracle.triplet_loss(pre.double(), goal.double(), post.double())
  assert abs(got.item() - float(want)) < 1e-4 * abs(float(want))
