"""CEM parity.  Golden vectors: tests/golden/cem_golden.npz, produced by the reference's own
utils/cross_entropy.py (tests/golden/make_cem_golden.py).

CPU: the oracle restatement and the host API mirror reproduce the golden runs exactly.
GPU: the elite-refit kernel reproduces every recorded iteration (selection is exact, mean/std to
fp32 rounding), the Philox sampling kernel matches the numpy restatement, and the on-device CEM +
Bellman target satisfy their invariants.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'cem_golden.npz'))
CASES = [0, 1, 2]


def _replay(cem_fn, update_fn, case):
  p = 'case%d_' % case
  noise, target, num_elites = GOLDEN[p + 'noise'], GOLDEN[p + 'target'], int(GOLDEN[p + 'num_elites'])
  it = {'i': 0}
  means, stds = [], []

  def sample_fn(mean, stddev):
    s = mean + stddev * noise[it['i']]
    it['i'] += 1
    return s

  def objective_fn(samples):
    return np.round(-np.sum((np.asarray(samples) - target)**2, axis=1), 1)

  def upd(params, elites):
    out = update_fn(params, elites)
    means.append(out['mean'])
    stds.append(out['stddev'])
    return out

  dim = noise.shape[2]
  samples, values, _ = cem_fn(sample_fn, objective_fn, upd, {'mean': np.zeros(dim), 'stddev': np.ones(dim)},
                              num_elites, num_iterations=noise.shape[0])
  return np.asarray(samples), np.asarray(values), np.stack(means), np.stack(stds)


@pytest.mark.parametrize('case', CASES)
def test_oracle_matches_reference_golden(case):
  from oracle import cem as oracle
  samples, values, means, stds = _replay(oracle.cross_entropy_method, oracle.normal_update_fn, case)
  p = 'case%d_' % case
  np.testing.assert_array_equal(means, GOLDEN[p + 'mean'])
  np.testing.assert_array_equal(stds, GOLDEN[p + 'stddev'])
  np.testing.assert_array_equal(samples, GOLDEN[p + 'samples'][-1])
  assert int(np.argmax(values)) == int(GOLDEN[p + 'best_index'])
  # batched form used to check the kernel
  for i in range(means.shape[0]):
    m, s, best, arg = oracle.refit_rows(GOLDEN[p + 'samples'][i][None], GOLDEN[p + 'values'][i][None],
                                        int(GOLDEN[p + 'num_elites']))
    np.testing.assert_allclose(m[0], GOLDEN[p + 'mean'][i], rtol=0, atol=1e-12)
    np.testing.assert_allclose(s[0], GOLDEN[p + 'stddev'][i], rtol=0, atol=1e-12)


@pytest.mark.parametrize('case', CASES)
def test_host_api_matches_reference_golden(case):
  from tensor2robot_b200.utils import cross_entropy

  def update_fn(params, elites):
    del params
    return {'mean': np.mean(elites, axis=0), 'stddev': np.std(elites, axis=0, ddof=1)}

  samples, values, means, stds = _replay(cross_entropy.CrossEntropyMethod, update_fn, case)
  p = 'case%d_' % case
  np.testing.assert_array_equal(means, GOLDEN[p + 'mean'])
  np.testing.assert_array_equal(stds, GOLDEN[p + 'stddev'])
  np.testing.assert_array_equal(samples[int(np.argmax(values))], GOLDEN[p + 'best_action'])


def test_normal_cem_matches_reference_golden():
  from tensor2robot_b200.utils import cross_entropy
  np.random.seed(123)
  mean, std = cross_entropy.NormalCrossEntropyMethod(lambda s: -np.sum((s - 0.5)**2, axis=1), np.zeros(4),
                                                     np.ones(4), 32, 6, 3)
  np.testing.assert_array_equal(mean, GOLDEN['normal_mean'])
  np.testing.assert_array_equal(std, GOLDEN['normal_stddev'])


def test_dict_sample_batches_and_threshold():
  from tensor2robot_b200.utils import cross_entropy
  calls = {'n': 0}

  def sample_fn(mean):
    calls['n'] += 1
    return {'a': [mean + i for i in range(5)], 'b': [10 * i for i in range(5)]}

  def update_fn(params, elites):
    assert elites['a'] == [params['mean'] + 3, params['mean'] + 4] and elites['b'] == [30, 40]
    return {'mean': params['mean'] + 1}

  _, values, params = cross_entropy.CrossEntropyMethod(sample_fn, lambda s: list(s['a']), update_fn, {'mean': 0}, 2,
                                                       num_iterations=5, threshold_to_terminate=5.5)
  assert calls['n'] == 3 and params == {'mean': 3} and max(values) == 6   # stops once max(values) > threshold


# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
def test_refit_kernel_matches_reference_golden(case):
  from tensor2robot_b200 import _lib
  p = 'case%d_' % case
  samples, values = GOLDEN[p + 'samples'], GOLDEN[p + 'values']       # [iters, A, D], [iters, A]
  iters, a, d = samples.shape
  e = int(GOLDEN[p + 'num_elites'])
  s = torch.from_numpy(samples.astype(np.float32)).cuda()
  v = torch.from_numpy(values.astype(np.float32)).cuda()
  mean = torch.empty((iters, d), device='cuda')
  std = torch.empty((iters, d), device='cuda')
  best = torch.empty(iters, device='cuda')
  idx = torch.empty(iters, dtype=torch.int32, device='cuda')
  ptr = lambda t: C.c_void_p(t.data_ptr())
  _lib.call('t2r_cem_refit', ptr(s), ptr(v), ptr(mean), ptr(std), ptr(best), ptr(idx), iters, a, d, e, None)
  torch.cuda.synchronize()
  # elite selection is exact, so mean/std only differ by fp32 rounding of the inputs and sums
  np.testing.assert_allclose(mean.cpu().numpy(), GOLDEN[p + 'mean'], rtol=2e-6, atol=2e-6)
  np.testing.assert_allclose(std.cpu().numpy(), GOLDEN[p + 'stddev'], rtol=2e-5, atol=2e-6)
  np.testing.assert_array_equal(idx.cpu().numpy(), values.astype(np.float32).argmax(1))
  assert int(idx[-1]) == int(GOLDEN[p + 'best_index'])


@pytest.mark.gpu
def test_sample_kernel_matches_philox_oracle():
  from oracle import cem as oracle
  from tensor2robot_b200 import _lib
  b, a, d = 5, 64, 10
  rng = np.random.RandomState(0)
  mean = rng.uniform(-1, 1, (b, d)).astype(np.float32)
  std = rng.uniform(0.1, 2, (b, d)).astype(np.float32)
  out = torch.empty((b, a, d), device='cuda')
  ptr = lambda t: C.c_void_p(t.data_ptr())
  m_d, s_d = torch.from_numpy(mean).cuda(), torch.from_numpy(std).cuda()
  _lib.call('t2r_cem_sample', ptr(m_d), ptr(s_d), ptr(out), b, a, d, 1234, 7, None)
  torch.cuda.synchronize()
  ref = oracle.cem_sample(mean, std, a, 1234, 7)
  # identical integer stream; logf/sincosf differ from numpy by a few ulp
  np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=2e-5, atol=2e-5)
  z = (out.cpu().numpy() - mean[:, None]) / std[:, None]
  assert abs(z.mean()) < 0.08 and abs(z.std() - 1) < 0.05


@pytest.mark.gpu
def test_device_cem_and_bellman_target_invariants():
  from tensor2robot_b200 import engine, nn
  from tensor2robot_b200.models import optimizers
  from tensor2robot_b200.research.qtopt import networks
  critic = networks.Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom()
  step = engine.CriticTrainStep(critic, optimizers.MomentumOptimizer(1e-4), device='cuda', seed=0)
  b = 3
  g = torch.Generator(device='cuda').manual_seed(0)
  frames = torch.randint(0, 256, (b, 512, 640, 3), dtype=torch.uint8, device='cuda', generator=g)
  step.build(frames, torch.zeros((b, 10), device='cuda'))
  cem = engine.CEMTargetComputer(critic, step.vs, action_size=10, cem_samples=64, cem_iters=2, num_elites=10, seed=3)
  x = step.preprocess(frames, training=False)
  action, max_q, dbg = cem.maximize(x)
  q, samples = dbg['q'], dbg['samples']
  assert action.shape == (b, 10) and max_q.shape == (b,)
  # the returned action is the arg-max over the LAST iteration's samples (policies.py:162)
  np.testing.assert_array_equal(max_q.cpu().numpy(), q.max(1).values.cpu().numpy())
  arg = q.argmax(1)
  np.testing.assert_array_equal(action.cpu().numpy(), samples[torch.arange(b), arg].cpu().numpy())
  # Q of the chosen action, recomputed from scratch, is the reported maximum
  q_again = step.predict(frames, action)
  np.testing.assert_allclose(q_again.float().cpu().numpy().reshape(-1), max_q.cpu().numpy(), atol=2e-3)
  reward = torch.tensor([0.0, 1.0, 1.0], device='cuda')
  done = torch.tensor([0.0, 0.0, 1.0], device='cuda')
  y = cem.bellman_target(reward, done, max_q, gamma=0.9).cpu().numpy()
  mq = max_q.cpu().numpy()
  np.testing.assert_allclose(y, [0.9 * mq[0], 1 + 0.9 * mq[1], 1.0], rtol=1e-6)
  assert (y >= 0).all() and (y <= 1.9 + 1e-6).all()


@pytest.mark.gpu
def test_lagged_target_network_tracks_online_critic_with_a_lag():
  """engine.LaggedTarget (SURVEY A-23: theta' of the Bellman target): equals the online critic right after a
  refresh, stays frozen while the online critic trains, catches up at the next refresh; CEM evaluated on
  the target store is unaffected by online updates in between."""
  import torch
  from tensor2robot_b200 import engine
  from tensor2robot_b200.models import optimizers
  from tensor2robot_b200.research.qtopt import networks
  torch.manual_seed(0)
  critic = networks.Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom()
  step = engine.CriticTrainStep(critic, optimizers.MomentumOptimizer(learning_rate=0.05, momentum=0.9),
                                input_hw=(512, 640), target_hw=(472, 472))
  rng = np.random.RandomState(0)
  images = torch.from_numpy(rng.randint(0, 256, (4, 512, 640, 3)).astype(np.uint8)).cuda()
  actions = torch.from_numpy(rng.uniform(-1, 1, (4, 10)).astype(np.float32)).cuda()
  reward = torch.from_numpy((rng.uniform(size=(4, 1)) < 0.5).astype(np.float32)).cuda()
  step.build(images, actions)
  target = engine.LaggedTarget(step, update_every=3)
  target.build(images, actions)
  cem = engine.CEMTargetComputer(critic, target.vs, action_size=10, cem_samples=16, cem_iters=2, num_elites=4, seed=1)
  x = step.preprocess(images, training=False)
  assert torch.equal(target.vs.flat, step.vs.flat) and torch.equal(target.vs.state_flat, step.vs.state_flat)
  _, q0, _ = cem.maximize(x)
  step.step(images, actions, reward)
  step.step(images, actions, reward)
  assert not target.update()                                   # only 2 optimizer steps since the refresh
  assert not torch.equal(target.vs.flat, step.vs.flat)          # the online critic has moved on
  cem.calls = 0
  _, q1, _ = cem.maximize(x)
  assert torch.equal(q0, q1)                                    # the target is frozen in between
  step.step(images, actions, reward)
  assert target.update()                                        # third step: refresh
  assert torch.equal(target.vs.flat, step.vs.flat) and torch.equal(target.vs.state_flat, step.vs.state_flat)
  cem.calls = 0
  _, q2, _ = cem.maximize(x)
  assert not torch.equal(q0, q2)


@pytest.mark.gpu
def test_bellman_critic_train_step_is_cem_target_plus_supervised_step():
  """engine.BellmanCriticTrainStep (BASELINE config C3): the step trains on y = r + gamma (1 - done) max_a Q'(s', a).
  (1) terminal transitions (done = 1): y == r, and the step is the supervised step on r (same loss, same update);
  (2) non-terminal: y equals the stand-alone CEMTargetComputer on the same target store / counter, chunked CEM
  (cem_chunk = 1) gives the same targets, and r <= y <= r + gamma."""
  from tensor2robot_b200 import engine
  from tensor2robot_b200.models import optimizers
  from tensor2robot_b200.research.qtopt import networks
  rng = np.random.RandomState(3)
  b = 3
  images = torch.from_numpy(rng.randint(0, 256, (b, 512, 640, 3)).astype(np.uint8)).cuda()
  nxt = torch.from_numpy(rng.randint(0, 256, (b, 512, 640, 3)).astype(np.uint8)).cuda()
  actions = torch.from_numpy(rng.uniform(-1, 1, (b, 10)).astype(np.float32)).cuda()
  reward = torch.tensor([[0.0], [1.0], [0.0]], device='cuda')

  def make(cls, **kw):
    torch.manual_seed(0)
    np.random.seed(0)
    critic = networks.Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom()
    return cls(critic, optimizers.MomentumOptimizer(learning_rate=0.01, momentum=0.9), seed=0, **kw)

  plain = make(engine.CriticTrainStep)
  bell = make(engine.BellmanCriticTrainStep, gamma=0.9, cem_samples=16, cem_iters=2, num_elites=4)
  plain.build(images, actions)
  bell.build(images, actions)
  assert torch.equal(plain.vs.flat, bell.vs.flat)
  loss_plain = plain.step(images, actions, reward)
  loss_bell = bell.step(images, actions, reward, nxt, torch.ones((b, 1), device='cuda'))
  np.testing.assert_array_equal(bell.last_target.cpu().numpy(), reward.reshape(-1).cpu().numpy())
  # the same supervised step (up to the summation order of the fused BN statistics' atomics); the parameter updates of a
  # 3-frame batch of noise through 16 batch-normalised layers are too ill-conditioned to compare element-wise
  np.testing.assert_allclose(float(loss_plain), float(loss_bell), rtol=1e-3)
  assert plain.global_step == bell.global_step == 1

  chunked = make(engine.BellmanCriticTrainStep, gamma=0.9, cem_samples=16, cem_iters=2, num_elites=4, cem_chunk=1)
  chunked.build(images, actions)
  whole = make(engine.BellmanCriticTrainStep, gamma=0.9, cem_samples=16, cem_iters=2, num_elites=4)
  whole.build(images, actions)
  done = torch.zeros((b, 1), device='cuda')
  chunked.step(images, actions, reward, nxt, done)
  whole.step(images, actions, reward, nxt, done)
  y = whole.last_target.cpu().numpy()
  np.testing.assert_allclose(chunked.last_target.cpu().numpy(), y, atol=2e-3)   # batch-size dependent kernel paths
  r = reward.reshape(-1).cpu().numpy()
  assert (y >= r - 1e-6).all() and (y <= r + 0.9 + 1e-6).all() and (y > r).any()
  # the same target from the stand-alone computer on a fresh copy of the (still initial) target network
  ref = make(engine.BellmanCriticTrainStep, gamma=0.9, cem_samples=16, cem_iters=2, num_elites=4)
  ref.build(images, actions)
  _, max_q, _ = ref.cem.maximize(ref.preprocess(nxt, training=False))
  np.testing.assert_allclose(y, r + 0.9 * max_q.cpu().numpy(), rtol=1e-6)


@pytest.mark.gpu
def test_device_cem_high_precision_target():
  """CEMTargetComputer(high_precision=True): the staged tower and every Q batch run in nn.high_precision() on fp32
  frames.  Same sampler / refit, so with the same seed the two precisions draw the same first-iteration samples; the
  reported maximum is reproduced by a from-scratch high-precision PREDICT of the chosen action to 1e-4, and stays within
  the bf16 path's own error of the bf16 result."""
  from tensor2robot_b200 import engine
  from tensor2robot_b200.models import optimizers
  from tensor2robot_b200.research.qtopt import networks
  critic = networks.Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom()
  step = engine.CriticTrainStep(critic, optimizers.MomentumOptimizer(1e-4), device='cuda', seed=0)
  b = 2
  g = torch.Generator(device='cuda').manual_seed(1)
  frames = torch.randint(0, 256, (b, 512, 640, 3), dtype=torch.uint8, device='cuda', generator=g)
  step.build(frames, torch.zeros((b, 10), device='cuda'))
  fast = engine.CEMTargetComputer(critic, step.vs, cem_samples=32, cem_iters=2, num_elites=6, seed=5)
  exact = engine.CEMTargetComputer(critic, step.vs, cem_samples=32, cem_iters=2, num_elites=6, seed=5, high_precision=True)
  _, q_fast, dbg_fast = fast.maximize(step.preprocess(frames, training=False))
  x32 = step.preprocess(frames, training=False, out_dtype=torch.float32)
  assert x32.dtype == torch.float32
  action, q_exact, dbg = exact.maximize(x32)
  assert dbg['q'].dtype == torch.float32 and tuple(dbg['q'].shape) == (b, 32)
  again = step.predict(frames, action, high_precision=True).float().reshape(-1)
  np.testing.assert_allclose(again.cpu().numpy(), q_exact.cpu().numpy(), atol=1e-4)
  np.testing.assert_allclose(q_exact.cpu().numpy(), q_fast.cpu().numpy(), atol=5e-3)
  # and inside the Bellman step
  bell = engine.BellmanCriticTrainStep(networks.Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom(),
                                       optimizers.MomentumOptimizer(1e-4), cem_samples=16, cem_iters=2, num_elites=4,
                                       high_precision_target=True, seed=0)
  actions = torch.zeros((b, 10), device='cuda')
  reward = torch.tensor([[0.0], [1.0]], device='cuda')
  loss = bell.step(frames, actions, reward, frames, torch.zeros((b, 1), device='cuda'))
  y = bell.last_target.cpu().numpy()
  assert np.isfinite(float(loss)) and (y >= reward.reshape(-1).cpu().numpy() - 1e-6).all() and (y <= 1.9 + 1e-6).all()
