"""pose_env (BASELINE config C1, the reference's own CPU-runnable case) on the GPU: the fp32 small-network kernels
(csrc/vision_small.cu) against the float64 oracle (oracle/vision_layers.py) - forward values and every gradient -
and both T2R models trained on the reference's fixture test_data/pose_env_test_data.tfrecord
(research/pose_env/pose_env_models_test.py:74-87: test_mc, test_regression)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, 'golden', 'pose_env_test_data.tfrecord')


def _rel(a, b):
  a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
  return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _randomise(vs, seed):
  """Moves every parameter off its initial value (gamma 1 / beta 0 / constant biases hide indexing errors)."""
  rng = np.random.RandomState(seed)
  arrays = vs.export_tf()
  for k, a in arrays.items():
    if k.endswith('gamma'):
      arrays[k] = (1.0 + 0.3 * rng.standard_normal(a.shape)).astype(np.float32)
    elif k.endswith(('beta', 'biases')):
      arrays[k] = (0.2 * rng.standard_normal(a.shape)).astype(np.float32)
    else:
      arrays[k] = (a + 0.05 * rng.standard_normal(a.shape)).astype(np.float32)
  vs.import_tf(arrays)
  return {k: torch.from_numpy(v.astype(np.float64)) for k, v in arrays.items()}


def _check_grads(grads, ov, tol=2e-4):
  assert sorted(grads) == sorted(ov)
  worst = 0.0
  for k, g in grads.items():
    go = ov[k].grad.numpy()
    assert np.isfinite(g).all(), k
    err = _rel(g, go)
    worst = max(worst, err)
    assert err < tol, (k, err, float(np.abs(go).max()))
  return worst


@pytest.mark.parametrize('batch', [1, 5])
def test_regression_network_matches_oracle(batch):
  from oracle import vision_layers as oracle
  from tensor2robot_b200 import nn
  from tensor2robot_b200.research.pose_env import pose_env_models as pm
  from tensor2robot_b200.utils import tensorspec_utils as tu
  rng = np.random.RandomState(3 + batch)
  img = rng.uniform(0, 1, (batch, 64, 64, 3)).astype(np.float32)
  target = rng.uniform(-1, 1, (batch, 2)).astype(np.float32)
  reward = (rng.uniform(size=(batch, 1)) < 0.7).astype(np.float32)
  reward[0, 0] = 1.0
  model = pm.PoseEnvRegressionModel()
  vs = nn.VariableStore('cuda:0', seed=11)
  feats = tu.TensorSpecStruct(state=torch.from_numpy(img).cuda())
  labels = tu.TensorSpecStruct(target_pose=torch.from_numpy(target).cuda(), reward=torch.from_numpy(reward).cuda())

  def run():
    out = model.a_func(feats, 'a_func', 'train')
    return out, model.loss_fn(labels, out, 'train')

  with torch.no_grad(), nn.variable_store(vs):
    run()
  vs.finalize()
  names = sorted(vs.export_tf())
  assert 'a_func/state_features/conv2/weights' in names and 'a_func/state_features/conv2/LayerNorm/gamma' in names
  assert 'a_func/state_features/final_conv_1x1/weights' in names and 'a_func/BiasAdd/biases' in names
  assert 'a_func/pose_fc0/LayerNorm/beta' in names and 'a_func/pose_fc2/biases' in names
  assert not any(n.endswith('conv2/biases') or n.endswith('pose_fc0/biases') for n in names)   # normalised: no bias
  assert vs.export_tf()['a_func/state_features/conv2/weights'].shape == (3, 3, 3, 32)
  assert vs.export_tf()['a_func/pose_fc0/weights'].shape == (74, 100)
  ov = _randomise(vs, 5)
  with nn.variable_store(vs):
    out, loss = run()
    vs.zero_grad()
    loss.backward()
  torch.cuda.synchronize()
  for v in ov.values():
    v.requires_grad_(True)
  pose_o, points_o = oracle.regression_a_func(torch.from_numpy(img.astype(np.float64)), ov)
  loss_o = oracle.weighted_mse(torch.from_numpy(target.astype(np.float64)), pose_o, torch.from_numpy(reward.astype(np.float64)))
  loss_o.backward()
  assert tuple(out['state_features'].shape) == (batch, 64) and tuple(out['inference_output'].shape) == (batch, 2)
  e_points = _rel(out['state_features'].detach().cpu().numpy(), points_o.detach().numpy())
  e_pose = _rel(out['inference_output'].detach().cpu().numpy(), pose_o.detach().numpy())
  e_loss = abs(float(loss) - float(loss_o)) / abs(float(loss_o))
  worst = _check_grads(vs.export_tf_grads(), ov)
  print('regression B=%d: points %.2e pose %.2e loss %.2e worst grad %.2e' % (batch, e_points, e_pose, e_loss, worst))
  assert e_points < 1e-5 and e_pose < 1e-5 and e_loss < 1e-5


@pytest.mark.parametrize('batch,actions', [(4, 4), (1, 6)])
def test_mc_critic_matches_oracle(batch, actions):
  """actions > batch exercises the tf.tile branch of the action merge (pose_env_models.py:141-149)."""
  from oracle import vision_layers as oracle
  from tensor2robot_b200 import nn
  from tensor2robot_b200.research.pose_env import pose_env_models as pm
  from tensor2robot_b200.utils import tensorspec_utils as tu
  rng = np.random.RandomState(7 + actions)
  img = rng.uniform(0, 1, (batch, 64, 64, 3)).astype(np.float32)
  pose = rng.uniform(-1, 1, (actions, 2)).astype(np.float32)
  reward = rng.uniform(-1, 0, (actions,)).astype(np.float32)
  model = pm.PoseEnvContinuousMCModel()
  vs = nn.VariableStore('cuda:0', seed=12)
  feats = tu.TensorSpecStruct()
  feats['state/image'] = torch.from_numpy(img).cuda()
  feats['action/pose'] = torch.from_numpy(pose).cuda()
  labels = tu.TensorSpecStruct(reward=torch.from_numpy(reward).cuda())

  def run():
    out = model.q_func(feats, 'q_func', 'train')
    return out, model.loss_fn(feats, labels, out)

  with torch.no_grad(), nn.variable_store(vs):
    run()
  vs.finalize()
  names = sorted(vs.export_tf())
  for n in ('q_func/q_features/Conv/weights', 'q_func/q_features/Conv_2/LayerNorm/gamma',
            'q_func/q_features/fully_connected/biases', 'q_func/Stack/fully_connected_2/weights',
            'q_func/fully_connected/weights'):
    assert n in names, (n, names)
  assert vs.export_tf()['q_func/Stack/fully_connected_1/weights'].shape == (7 * 7 * 32, 100)
  ov = _randomise(vs, 6)
  with nn.variable_store(vs):
    out, loss = run()
    vs.zero_grad()
    loss.backward()
  torch.cuda.synchronize()
  for v in ov.values():
    v.requires_grad_(True)
  q_o = oracle.mc_critic_q(torch.from_numpy(img.astype(np.float64)), torch.from_numpy(pose.astype(np.float64)), ov)
  loss_o = oracle.weighted_mse(torch.from_numpy(reward.astype(np.float64)), q_o)
  loss_o.backward()
  assert tuple(out['q_predicted'].shape) == (actions,)
  e_q = _rel(out['q_predicted'].detach().cpu().numpy(), q_o.detach().numpy())
  worst = _check_grads(vs.export_tf_grads(), ov)
  print('mc critic B=%d A=%d: q %.2e worst grad %.2e' % (batch, actions, e_q, worst))
  assert e_q < 1e-5


def test_small_kernels_edge_shapes():
  """Direct convolution with SAME padding / stride 2 / odd channel counts, and a layer norm over rows that are not a
  multiple of the block size, against torch float64 on the host."""
  import torch.nn.functional as F
  from tensor2robot_b200 import nn
  rng = np.random.RandomState(0)
  x = rng.standard_normal((3, 9, 11, 5)).astype(np.float32)
  vs = nn.VariableStore('cuda:0', seed=2)
  xg = torch.from_numpy(x).cuda().requires_grad_(True)

  def run():
    y = nn.conv2d_f32(xg, 7, (3, 5), stride=2, padding='SAME', use_bias=True, scope='c', bias_initializer=0.1)
    return nn.layer_norm(y, scope='ln', relu=False)

  with torch.no_grad(), nn.variable_store(vs):
    run()
  vs.finalize()
  ov = _randomise(vs, 1)
  t = rng.standard_normal((3, 5, 6, 7)).astype(np.float32)
  with nn.variable_store(vs):
    y = run()
    vs.zero_grad()
    (y * torch.from_numpy(t).cuda()).sum().backward()
  torch.cuda.synchronize()
  assert tuple(y.shape) == (3, 5, 6, 7)
  for v in ov.values():
    v.requires_grad_(True)
  xo = torch.from_numpy(x.astype(np.float64)).requires_grad_(True)
  # TF SAME for H=9,k=3,s=2: out 5, pad total 2 -> (1,1); W=11,k=5,s=2: out 6, pad total 4 -> (2,2)
  xp = F.pad(xo.permute(0, 3, 1, 2), (2, 2, 1, 1))
  yo = F.conv2d(xp, ov['c/weights'].permute(3, 2, 0, 1), ov['c/biases'], stride=2).permute(0, 2, 3, 1)
  mean = yo.mean((1, 2, 3), keepdim=True)
  var = ((yo - mean) ** 2).mean((1, 2, 3), keepdim=True)
  yo = (yo - mean) / torch.sqrt(var + 1e-12) * ov['ln/gamma'] + ov['ln/beta']
  (yo * torch.from_numpy(t.astype(np.float64))).sum().backward()
  assert _rel(y.detach().cpu().numpy(), yo.detach().numpy()) < 1e-5
  assert _rel(xg.grad.cpu().numpy(), xo.grad.numpy()) < 1e-4
  _check_grads(vs.export_tf_grads(), ov)


@pytest.mark.parametrize('which', ['regression', 'mc'])
def test_train_on_reference_fixture(tmp_path, which):
  """pose_env_models_test.py:74-87 with the reference's own 100-record fixture: 3 train steps + 2 eval steps."""
  from tensor2robot_b200.input_generators import default_input_generator as gens
  from tensor2robot_b200.research.pose_env import pose_env_models as pm
  from tensor2robot_b200.utils import train_eval
  model = pm.PoseEnvRegressionModel() if which == 'regression' else pm.PoseEnvContinuousMCModel()
  out = train_eval.train_eval_model(
      t2r_model=model, input_generator_train=gens.DefaultRecordInputGenerator(batch_size=2, file_patterns=FIXTURE),
      input_generator_eval=gens.DefaultRecordInputGenerator(batch_size=2, file_patterns=FIXTURE),
      max_train_steps=3, eval_steps=2, model_dir=str(tmp_path))
  assert out['global_step'] == 3 and np.isfinite(out['loss'])
  state = torch.load(str(tmp_path / 'model.ckpt-3.pt'), weights_only=False)
  init = torch.load(str(tmp_path / 'model.ckpt-0.pt'), weights_only=False)
  moved = [k for k in init['variables'] if np.abs(state['variables'][k] - init['variables'][k]).max() > 0]
  assert len(moved) == len(init['variables']), sorted(set(init['variables']) - set(moved))


@pytest.mark.parametrize('normalizer,with_film', [('layer_norm', True), ('batch_norm', False), ('batch_norm', True)])
def test_vision_tower_film_and_batch_norm_variants_match_oracle(normalizer, with_film):
  """layers/vision_layers.py:72-86 (slim.batch_norm normaliser: decay .99, eps 1e-4, scale only on the final 1x1
  convolution) and :100-141, 162-181 (FiLM: (1 + gamma) * h + beta before the ReLU, parameters from a linear layer on
  the embedding): feature points, every gradient and the moving statistics against the float64 oracle."""
  from oracle import vision_layers as oracle
  from tensor2robot_b200 import nn
  from tensor2robot_b200.layers import vision_layers
  rng = np.random.RandomState(17)
  b = 6
  img = rng.uniform(0, 1, (b, 64, 64, 3)).astype(np.float32)
  emb = rng.standard_normal((b, 12)).astype(np.float32)
  wts = rng.standard_normal((b, 64)).astype(np.float32)
  vs = nn.VariableStore('cuda:0', seed=4)
  img_t, emb_t, wts_t = (torch.from_numpy(a).cuda() for a in (img, emb, wts))

  def run(training):
    with nn.variable_scope('tower'):
      film = vision_layers.BuildFILMParams(emb_t) if with_film else None
      points, extra = vision_layers.BuildImagesToFeaturesModel(img_t, is_training=training, normalizer_fn=normalizer,
                                                               film_output_params=film)
    return points, extra

  with torch.no_grad(), nn.variable_store(vs):
    run(False)
  vs.finalize()
  names = sorted(vs.export_tf())
  tag = 'LayerNorm' if normalizer == 'layer_norm' else 'BatchNorm'
  assert ('tower/conv2/%s/beta' % tag) in names and ('tower/final_conv_1x1/%s/gamma' % tag) in names
  if normalizer == 'batch_norm':
    assert 'tower/conv2/BatchNorm/gamma' not in names and 'tower/conv2/BatchNorm/moving_variance' in names   # scale=False
  if with_film:
    assert vs.export_tf()['tower/film/weights'].shape == (12, 2 * 5 * 32)
  ov = _randomise(vs, 9)
  for k in ov:
    if k.endswith('moving_variance'):
      ov[k] = ov[k].abs() + 0.5
  vs.import_tf({k: v.numpy().astype(np.float32) for k, v in ov.items()})
  for v in ov.values():
    v.requires_grad_(True)
  with nn.variable_store(vs):
    points, extra = run(True)
    loss = (points * wts_t).sum()
    vs.zero_grad()
    loss.backward()
  torch.cuda.synchronize()
  assert tuple(points.shape) == (b, 64) and tuple(extra['softmax'].shape)[0] == b
  updates = {}
  film_o = None
  if with_film:
    film_o = torch.from_numpy(emb.astype(np.float64)) @ ov['tower/film/weights'] + ov['tower/film/biases'].reshape(1, -1)
  points_o = oracle.images_to_features(torch.from_numpy(img.astype(np.float64)), ov, 'tower', film=film_o,
                                       normalizer=normalizer, training=True, updates=updates)
  (points_o * torch.from_numpy(wts.astype(np.float64))).sum().backward()
  err = _rel(points.detach().cpu().numpy(), points_o.detach().numpy())
  grads = vs.export_tf_grads()
  worst = 0.0
  for k, g in grads.items():
    go = ov[k].grad
    if go is None:
      continue
    worst = max(worst, _rel(g, go.numpy()))
  new = vs.export_tf()
  for k, u in updates.items():
    assert _rel(new[k], u.detach().numpy()) < 1e-5, k
  print('%s film=%s: points rel %.2e, worst gradient rel %.2e, %d moving statistics' % (normalizer, with_film, err, worst,
                                                                                       len(updates)))
  assert err < 2e-5 and worst < 5e-4
  # inference mode reads the moving statistics
  with torch.no_grad(), nn.variable_store(vs):
    p_eval, _ = run(False)
  ov2 = {k: torch.from_numpy(np.asarray(v, np.float64)) for k, v in new.items()}
  film_e = None
  if with_film:
    film_e = torch.from_numpy(emb.astype(np.float64)) @ ov2['tower/film/weights'] + ov2['tower/film/biases'].reshape(1, -1)
  p_eval_o = oracle.images_to_features(torch.from_numpy(img.astype(np.float64)), ov2, 'tower', film=film_e,
                                       normalizer=normalizer, training=False)
  assert _rel(p_eval.cpu().numpy(), p_eval_o.numpy()) < 2e-5
