"""GPU parity of the ResNet v2 tower (layers/film_resnet_model.py) and of the ResNet-50 Q-critic
composition against the oracle restatement (oracle/resnet.py), on identical weights / inputs.

Inference mode (moving statistics) is well conditioned and asserted tightly against the
bf16-storage oracle.  The training step uses the conditioning-relative criterion explained in
tests/test_qtopt_networks_gpu.py (random-init BN-ReLU nets amplify rounding differences)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
COND_FACTOR, GRAD_FLOOR = 1.5, 0.05


def _rel_l2(a, b):
  a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
  return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12))


def _images(b, size, seed):
  rng = np.random.RandomState(seed)
  yy, xx = np.mgrid[0:size, 0:size].astype(np.float32) / size
  img = np.zeros((b, size, size, 3), np.float32)
  for i in range(b):
    img[i] = rng.uniform(0.1, 0.9, size=(1, 1, 3))
    for _ in range(5):
      cy, cx, r = rng.uniform(0, 1, 3) * [1, 1, 0.3] + [0, 0, 0.05]
      m = np.exp(-((yy - cy)**2 + (xx - cx)**2) / (2 * r * r))[..., None]
      img[i] = img[i] * (1 - m) + rng.uniform(0, 1, 3) * m
  return np.clip(img + rng.uniform(-.03, .03, img.shape), 0, 1).astype(np.float32)


def _oracle_variables(img, grasp, resnet_size, seed):
  from oracle import resnet as oracle
  variables = {}
  with torch.no_grad():
    oracle.critic(variables, torch.from_numpy(img[:2]), torch.from_numpy(grasp[:2]), False, resnet_size=resnet_size,
                  rng=np.random.RandomState(seed))
  rng = np.random.RandomState(seed + 1)
  for k in list(variables):
    v = variables[k]
    if k.endswith('gamma'):
      variables[k] = (1 + 0.2 * torch.from_numpy(rng.randn(*v.shape).astype(np.float32)))
    elif k.endswith('beta'):
      variables[k] = 0.1 * torch.from_numpy(rng.randn(*v.shape).astype(np.float32))
    elif k.endswith('moving_variance'):
      variables[k] = torch.from_numpy(rng.uniform(0.5, 1.5, v.shape).astype(np.float32))
    elif k.endswith('moving_mean'):
      variables[k] = 0.1 * torch.from_numpy(rng.randn(*v.shape).astype(np.float32))
    elif k.endswith('/weights'):
      variables[k] = v * 8.0
  return variables


def _engine(img_t, grasp_t, variables, resnet_size):
  from tensor2robot_b200 import nn
  from tensor2robot_b200.research.qtopt import resnet_critic
  vs = nn.VariableStore('cuda', seed=2)
  net = resnet_critic.ResNet50QCritic(resnet_size=resnet_size)
  with torch.no_grad(), nn.variable_store(vs):
    net.model((None, img_t[:2]), grasp_t[:2], is_training=False)
  vs.finalize()
  assert sorted(vs.export_tf().keys()) == sorted(variables.keys())      # identical reference variable names
  vs.import_tf({k: v.numpy() for k, v in variables.items()})
  return vs, net


@pytest.mark.parametrize('resnet_size', [50, 18])
def test_resnet_critic_inference_matches_oracle(resnet_size):
  from oracle import resnet as oracle, tf_ops
  from tensor2robot_b200 import nn
  b, a, size = 2, 8, 96
  img = _images(b, size, 0)
  grasp = np.random.RandomState(1).uniform(-1, 1, (b, a, 10)).astype(np.float32)
  variables = _oracle_variables(img, grasp[:, 0], resnet_size, 3)
  img_t = torch.from_numpy(img).cuda().to(torch.bfloat16)
  grasp_t = torch.from_numpy(grasp).cuda()
  vs, net = _engine(img_t, grasp_t[:, 0], variables, resnet_size)
  with torch.no_grad(), nn.variable_store(vs):
    logits, ep = net.model((None, img_t), grasp_t, is_training=False)
  q = ep['predictions'].float().cpu().numpy()
  assert q.shape == (b, a)
  res = {}
  for name, storage in (('bf16', torch.bfloat16), ('fp32', None)):
    tf_ops.STORAGE_DTYPE = storage
    try:
      ep_o = {}
      with torch.no_grad():
        lo = oracle.critic(dict(variables), img_t.float().cpu(), torch.from_numpy(grasp), False, resnet_size=resnet_size,
                           end_points=ep_o)
      res[name] = (lo.numpy().reshape(-1), ep_o['predictions'].numpy())
    finally:
      tf_ops.STORAGE_DTYPE = None
  le = logits.float().cpu().numpy().reshape(-1)
  print('resnet%d logits engine %s\n         bf16 oracle %s' % (resnet_size, le[:4], res['bf16'][0][:4]))
  print('rel_l2(logits) vs bf16-storage oracle %.3e, vs fp32 oracle %.3e; max|dq| %.3e / %.3e' % (
      _rel_l2(le, res['bf16'][0]), _rel_l2(le, res['fp32'][0]), np.abs(q - res['bf16'][1]).max(),
      np.abs(q - res['fp32'][1]).max()))
  # 53 bf16 layers compound to ~2e-2 on the logits of ResNet-50 (7e-3 for ResNet-18); with inference batch
  # norm folded into the convolutions the engine rounds at different points than the bf16-storage oracle
  assert _rel_l2(le, res['bf16'][0]) < (3e-2 if resnet_size == 50 else 1.5e-2)
  assert np.abs(q - res['bf16'][1]).max() < 1e-2
  assert np.abs(q - res['fp32'][1]).max() < 2e-2


def test_resnet50_critic_train_step_matches_oracle():
  from oracle import resnet as oracle, tf_ops
  from tensor2robot_b200 import nn
  b, size = 8, 96
  img = _images(b, size, 5)
  rng = np.random.RandomState(6)
  grasp = rng.uniform(-1, 1, (b, 10)).astype(np.float32)
  reward = (rng.uniform(size=(b, 1)) < 0.4).astype(np.float32)
  variables = _oracle_variables(img, grasp, 50, 7)
  img_t = torch.from_numpy(img).cuda().to(torch.bfloat16)
  grasp_t, reward_t = torch.from_numpy(grasp).cuda(), torch.from_numpy(reward).cuda()
  vs, net = _engine(img_t, grasp_t, variables, 50)
  with nn.variable_store(vs):
    logits, _ = net.model((None, img_t), grasp_t, is_training=True)
    loss, q = nn.sigmoid_log_loss(logits, reward_t)
    vs.zero_grad()
    loss.backward()
  torch.cuda.synchronize()
  grads = vs.export_tf_grads()
  new_vars = vs.export_tf()

  def oracle_step(storage):
    tf_ops.STORAGE_DTYPE = storage
    try:
      ov = {k: v.clone().requires_grad_(not k.endswith(('moving_mean', 'moving_variance'))) for k, v in variables.items()}
      updates = {}
      lo = oracle.critic(ov, img_t.float().cpu(), torch.from_numpy(grasp), True, updates=updates)
      qo = torch.sigmoid(lo)
      l = tf_ops.log_loss(torch.from_numpy(reward), qo)
      l.backward()
    finally:
      tf_ops.STORAGE_DTYPE = None
    return ov, updates, qo.detach().numpy().reshape(-1), float(l.detach())

  ov, updates, q_b, loss_b = oracle_step(torch.bfloat16)
  ov_f, _, q_f, loss_f = oracle_step(None)
  q_e = q.float().cpu().numpy().reshape(-1)
  err_b, cond_q = np.abs(q_e - q_b).max(), np.abs(q_b - q_f).max()
  print('loss engine %.5f bf16-oracle %.5f fp32-oracle %.5f; max|dq| engine-vs-bf16 %.3e, bf16-vs-fp32 %.3e' % (
      float(loss.detach()), loss_b, loss_f, err_b, cond_q))
  failures, worst = [], 0.0
  gmax = max(float(v.grad.norm()) for v in ov.values() if v.grad is not None)
  for k, g in grads.items():
    go = ov[k].grad
    if go is None or float(go.norm()) < 1e-4 * gmax:
      continue
    # tensors whose true gradient nearly cancels (e.g. the logit bias = sum(q - y)/n) are measured
    # against 1 % of the largest gradient norm instead of their own tiny norm
    den = max(float(go.norm()), 1e-2 * gmax)
    e = float(np.linalg.norm(np.asarray(g, np.float64) - go.numpy())) / den
    cond = float((ov_f[k].grad - go).norm()) / den
    worst = max(worst, e)
    if not e < COND_FACTOR * cond + GRAD_FLOOR:
      failures.append((k, e, cond))
  print('worst gradient rel_l2 vs bf16-storage oracle: %.3e over %d tensors' % (worst, len(grads)))
  for k, u in updates.items():
    assert _rel_l2(new_vars[k], u.numpy()) < 3e-2, k
  assert err_b < COND_FACTOR * cond_q + 2e-3
  assert not failures, failures[:5]


def test_film_resnet18_matches_oracle():
  """BC-Z style FiLM-conditioned ResNet (layers/resnet.py:98-209 + film_resnet_model.py:108-115):
  linear_film_generator -> per-block (gamma, beta) -> second batch norm of every block.  Inference
  forward is asserted tightly; the training step (FiLM backward kernels, gradients into the
  generator) against torch autograd on the oracle with the loose conditioning bound."""
  from oracle import resnet as oracle, tf_ops
  from tensor2robot_b200 import nn
  from tensor2robot_b200.layers import resnet
  b, size, classes, emb_dim = 16, 64, 64, 64
  img = _images(b, size, 11)
  emb = np.random.RandomState(12).standard_normal((b, emb_dim)).astype(np.float32)
  img_t = torch.from_numpy(img).cuda().to(torch.bfloat16)
  emb_t = torch.from_numpy(emb).cuda()
  vs = nn.VariableStore('cuda', seed=4)

  def run(training):
    return resnet.resnet_model(img_t, training, classes, resnet_size=18,
                               film_generator_fn=resnet.linear_film_generator, film_generator_input=emb_t)

  with torch.no_grad(), nn.variable_store(vs):
    run(False)
  vs.finalize()
  variables = {k: torch.from_numpy(np.array(v)) for k, v in vs.export_tf().items()}
  rng = np.random.RandomState(13)
  for k in list(variables):   # make the FiLM path matter: non-trivial generator weights / BN parameters
    if k.startswith('film') and k.endswith('weights'):
      variables[k] = torch.from_numpy(0.2 * rng.standard_normal(tuple(variables[k].shape)).astype(np.float32))
    elif k.endswith('moving_variance'):
      variables[k] = torch.from_numpy(rng.uniform(0.5, 1.5, tuple(variables[k].shape)).astype(np.float32))
  vs.import_tf({k: v.numpy() for k, v in variables.items()})
  film_keys = sorted(k for k in variables if k.startswith('film') and k.endswith('weights'))
  assert len(film_keys) == 4

  def oracle_films(v):
    e = torch.from_numpy(emb).to(torch.bfloat16).float()
    films = []
    for i, nb in enumerate([2, 2, 2, 2]):
      w = v['film%d/weights' % i]
      out = e @ w.to(torch.bfloat16).float() + v['film%d/biases' % i]
      films.append(list(tf_ops._store(out).split(out.shape[1] // nb, dim=-1)))
    return films

  # ---- inference ----
  with torch.no_grad(), nn.variable_store(vs):
    le = run(False).float().cpu().numpy()
  tf_ops.STORAGE_DTYPE = torch.bfloat16
  try:
    with torch.no_grad():
      lo = oracle.resnet_model(dict(variables), img_t.float().cpu(), False, classes, 18,
                               films=oracle_films(variables)).numpy()
  finally:
    tf_ops.STORAGE_DTYPE = None
  print('film resnet18 inference rel_l2 %.3e' % _rel_l2(le, lo))
  assert _rel_l2(le, lo) < 2e-2

  # ---- training step ----
  target = torch.from_numpy(np.random.RandomState(14).standard_normal((b, classes)).astype(np.float32))
  with nn.variable_store(vs):
    logits = run(True)
    vs.zero_grad()
    (nn.to_f32(logits) * target.cuda()).sum().backward()
  torch.cuda.synchronize()
  grads = vs.export_tf_grads()
  ov = {k: v.clone().requires_grad_(not k.endswith(('moving_mean', 'moving_variance'))) for k, v in variables.items()}
  tf_ops.STORAGE_DTYPE = torch.bfloat16
  try:
    lo_t = oracle.resnet_model(ov, img_t.float().cpu(), True, classes, 18, films=oracle_films(ov), updates={})
    (lo_t * target).sum().backward()
  finally:
    tf_ops.STORAGE_DTYPE = None
  for k in film_keys + ['film0/biases']:
    g, go = grads[k], ov[k].grad.numpy()
    assert np.isfinite(g).all() and np.abs(g).max() > 0
    print('%-16s rel_l2 %.3e' % (k, _rel_l2(g, go)))
    assert _rel_l2(g, go) < 0.3


def test_bcz_resnet_film_network_trains():
  """research/bcz/model.py:245-285 + layers/bcz_networks.py:107-145: FiLM-ResNet tower (ResNet-18 here) with one
  MLP head per pose component, 3 waypoints (first from `action_trajectory`, the rest from
  `auxiliary_trajectory` heads behind a stop_gradient); every trainable variable except the unused
  classification head receives a gradient and the output shapes / variable scopes are the reference's."""
  from tensor2robot_b200 import nn
  from tensor2robot_b200.layers import resnet
  from tensor2robot_b200.research.bcz import model as bcz
  from tensor2robot_b200.utils import tensorspec_utils
  b = 4
  img = torch.from_numpy(_images(b, 96, 21)).cuda().to(torch.bfloat16)
  emb = torch.from_numpy(np.random.RandomState(22).standard_normal((b, 64)).astype(np.float32)).cuda()
  feats = tensorspec_utils.TensorSpecStruct(image=img)
  comps = [('xyz', 3, True, 100.), ('quaternion', 4, False, 10.), ('target_close', 1, False, 1.)]

  def run(mode):
    return bcz.resnet_film_network(feats, mode, comps, num_waypoints=3, film_generator_fn=resnet.linear_film_generator,
                                   condition_input=emb, resnet_size=18)

  vs = nn.VariableStore('cuda', seed=1)
  with torch.no_grad(), nn.variable_store(vs):
    run('eval')
  vs.finalize()
  names = list(vs.export_tf().keys())
  assert any(n.startswith('vision_model/action_trajectory/Stack/fully_connected_1/') for n in names)
  assert any(n.startswith('vision_model/auxiliary_trajectory/') for n in names)
  assert any(n.startswith('vision_model/film0/') for n in names)
  with nn.variable_store(vs):
    out, state = run('train')
    assert sorted(out) == ['policy_image_features', 'quaternion', 'target_close', 'xyz_residual']
    assert tuple(out['xyz_residual'].shape) == (b, 3, 3) and tuple(out['quaternion'].shape) == (b, 3, 4)
    assert tuple(out['policy_image_features'].shape) == (b, 512) and tuple(state.shape) == (b, 256)
    loss = sum((out[k] ** 2).sum() for k in ('xyz_residual', 'quaternion', 'target_close')) + (state ** 2).sum()
    vs.zero_grad()
    loss.backward()
  torch.cuda.synchronize()
  grads = vs.export_tf_grads()
  dead = [k for k, g in grads.items() if not np.abs(g).max() > 0 and '/dense/' not in k]
  assert np.isfinite(loss.item()) and not dead, dead[:5]


@pytest.mark.parametrize('resnet_size', [18, 50])
def test_resnet_v1_matches_oracle(resnet_size):
  """ResNet v1 (post-activation) blocks (layers/film_resnet_model.py:121-168, 220-276; BN + ReLU after the stem
  :565-571, projection shortcut with its own batch norm, shortcut add then ReLU, no final batch norm): inference
  logits against the bf16-storage oracle and one training step's gradients against it."""
  from oracle import resnet as oracle, tf_ops
  from tensor2robot_b200 import nn
  from tensor2robot_b200.layers import film_resnet_model as rm
  from tensor2robot_b200.layers import resnet as resnet_factory
  b, size, classes = 4, 64, 64
  img_t = torch.from_numpy(_images(b, size, 31)).cuda().to(torch.bfloat16)
  vs = nn.VariableStore('cuda', seed=6)
  model = rm.Model(resnet_size=resnet_size, bottleneck=resnet_size >= 50, num_classes=classes, num_filters=64,
                   kernel_size=7, conv_stride=2, first_pool_size=3, first_pool_stride=2,
                   block_sizes=resnet_factory._get_block_sizes(resnet_size), block_strides=[1, 2, 2, 2],
                   weight_decay=1e-4, resnet_version=1)
  with torch.no_grad(), nn.variable_store(vs):
    model(img_t, False)
  vs.finalize()
  variables = {k: torch.from_numpy(np.array(v)) for k, v in vs.export_tf().items()}
  rng = np.random.RandomState(32)
  for k in list(variables):
    if k.endswith('moving_variance'):
      variables[k] = torch.from_numpy(rng.uniform(0.5, 1.5, tuple(variables[k].shape)).astype(np.float32))
    elif k.endswith('moving_mean') or k.endswith('beta'):
      variables[k] = torch.from_numpy(0.1 * rng.standard_normal(tuple(variables[k].shape)).astype(np.float32))
  vs.import_tf({k: v.numpy() for k, v in variables.items()})
  n_bn = len([k for k in variables if k.endswith('/beta')])
  # v1: stem BN + 2 (3) per block + one per block layer (the first block of EVERY layer projects, film_resnet_model.py
  # :343-388): ResNet-18 1 + 2*8 + 4 = 21, ResNet-50 1 + 3*16 + 4 = 53
  assert n_bn == (21 if resnet_size == 18 else 53)
  with torch.no_grad(), nn.variable_store(vs):
    le = model(img_t, False).float().cpu().numpy()
  tf_ops.STORAGE_DTYPE = torch.bfloat16
  try:
    with torch.no_grad():
      lo = oracle.resnet_model(dict(variables), img_t.float().cpu(), False, classes, resnet_size, version=1).numpy()
  finally:
    tf_ops.STORAGE_DTYPE = None
  err = _rel_l2(le, lo)
  print('resnet%d v1 inference rel_l2 %.3e' % (resnet_size, err))
  assert err < 2e-2
  target = torch.from_numpy(np.random.RandomState(33).standard_normal((b, classes)).astype(np.float32))
  with nn.variable_store(vs):
    logits = model(img_t, True)
    vs.zero_grad()
    (nn.to_f32(logits) * target.cuda()).sum().backward()
  torch.cuda.synchronize()
  grads = vs.export_tf_grads()
  ov = {k: v.clone().requires_grad_(not k.endswith(('moving_mean', 'moving_variance'))) for k, v in variables.items()}
  tf_ops.STORAGE_DTYPE = torch.bfloat16
  try:
    lo_t = oracle.resnet_model(ov, img_t.float().cpu(), True, classes, resnet_size, updates={}, version=1)
    (lo_t * target).sum().backward()
  finally:
    tf_ops.STORAGE_DTYPE = None
  # training-mode gradients of a randomly initialised 18/50-layer BN network are ill conditioned (see the module doc):
  # the value gate is the inference comparison above; here every variable must receive a finite, non-zero gradient
  # that points the same way as the oracle's for the well-conditioned last layer
  k = 'resnet_model/dense/kernel'
  g, go = grads[k].reshape(-1), ov[k].grad.numpy().reshape(-1)
  cosine = float(np.dot(g, go) / (np.linalg.norm(g) * np.linalg.norm(go)))
  print('resnet%d v1 train: cosine(dense kernel gradient, oracle) %.4f' % (resnet_size, cosine))
  assert cosine > 0.9
  assert all(np.isfinite(v).all() for v in grads.values())
  assert all(np.abs(v).max() > 0 for v in grads.values())
