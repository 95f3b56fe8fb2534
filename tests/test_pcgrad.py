"""PCGrad.  CPU: the oracle against the known-answer vectors of research/qtopt/pcgrad_test.py:35-129 and
the Gram-matrix formulation the kernel uses against the sequential one.  GPU: `t2r_pcgrad_project` through the PCGrad
wrapper on the same known answers (all allow / deny list cases), on random multi-task gradients, and through
`compute_gradients` on a real network with two task losses."""
import numpy as np
import pytest
import torch

from oracle import pcgrad as oracle

# pcgrad_test.py:41-46: loss0 = var0.c0 + var1.c2, loss1 = var0.c1 + var1.c0
VAR0, VAR1 = np.array([1.0, 2.0]), np.array([3.0, 4.0])
C0, C1, C2 = np.array([1., 0.]), np.array([-1., -1.]), np.array([-1., 1.])
TASK_GRADS = [{'first_var/var0': C0, 'second_var/var1': C2}, {'first_var/var0': C1, 'second_var/var1': C0}]
NAMES = ['first_var/var0', 'second_var/var1']
KAT_PCGRAD = {'first_var/var0': [0.5, -1.5], 'second_var/var1': [0.5, 1.5]}         # pcgrad_test.py:89-97
KAT_RESULT = {'first_var/var0': [0.9995, 2.0015], 'second_var/var1': [2.9995, 3.9985]}
KAT_PLAIN = {'first_var/var0': [0.0, -1.0], 'second_var/var1': [0.0, 1.0]}          # pcgrad_test.py:100-103
CASES = [(None, None, [0, 1]), (None, ['*var*'], [0, 1]), (['second*'], None, [0]), (None, ['first*'], [0]),
         (None, ['*0'], [0]), (['first*'], None, [1]), (['*var*'], None, [])]      # (denylist, allowlist, pcgrad vars)


@pytest.mark.parametrize('denylist,allowlist,pcgrad_idx', CASES)
def test_oracle_known_answers(denylist, allowlist, pcgrad_idx):
  got = oracle.compute_gradients(TASK_GRADS, NAMES, allowlist=allowlist, denylist=denylist)
  for i, n in enumerate(NAMES):
    np.testing.assert_allclose(got[n], KAT_PCGRAD[n] if i in pcgrad_idx else KAT_PLAIN[n], atol=1e-5)
  for i in pcgrad_idx:     # SGD with lr 0.001 on the projected gradient (pcgrad_test.py:119-123)
    n = NAMES[i]
    np.testing.assert_allclose([VAR0, VAR1][i] - 0.001 * got[n], KAT_RESULT[n], atol=1e-6)


def _gram_formulation(task_grads, eps=1e-5):
  """What csrc/pcgrad.cu computes: the projections on coefficient vectors over the task gradients."""
  g = np.stack([np.asarray(x, np.float64).ravel() for x in task_grads])
  gram = g @ g.T
  t = len(g)
  a = np.zeros(t)
  for i in range(t):
    c = np.zeros(t)
    c[i] = 1
    for k in range(t):
      pd = (c @ gram[:, k]) / (gram[k, k] + eps)
      if pd < 0:
        c[k] -= pd
    a += c
  return (a @ g).reshape(np.shape(task_grads[0]))


@pytest.mark.parametrize('tasks', [1, 2, 3, 5, 8])
def test_gram_formulation_equals_sequential(tasks):
  rng = np.random.RandomState(tasks)
  for _ in range(20):
    gs = [rng.standard_normal((4, 3)) for _ in range(tasks)]
    if tasks > 2:
      gs[1] = np.zeros((4, 3))          # a task that does not touch the variable (None gradient in the reference)
    np.testing.assert_allclose(_gram_formulation(gs), oracle.project_variable(gs), rtol=1e-9, atol=1e-12)


def test_variable_selection():
  from tensor2robot_b200.research.qtopt import pcgrad
  for denylist, allowlist, idx in CASES:
    opt = pcgrad.PCGrad(None, allowlist=allowlist, denylist=denylist)
    assert [i for i, n in enumerate(NAMES) if opt.uses_pcgrad(n)] == idx
    assert [i for i, n in enumerate(NAMES) if oracle.uses_pcgrad(n, allowlist, denylist)] == idx
  with pytest.raises(AssertionError):
    pcgrad.PCGrad(None).compute_gradients(torch.zeros(()), None)


def _kat_store():
  from tensor2robot_b200 import nn
  vs = nn.VariableStore('cuda:0', seed=0)
  with nn.variable_store(vs):
    vs.get_variable('first_var/var0', (2,), lambda shape, rng: VAR0.astype(np.float32))
    vs.get_variable('second_var/var1', (2,), lambda shape, rng: VAR1.astype(np.float32))
  vs.finalize()
  return vs


@pytest.mark.gpu
@pytest.mark.parametrize('denylist,allowlist,pcgrad_idx', CASES)
def test_kernel_known_answers(denylist, allowlist, pcgrad_idx):
  from tensor2robot_b200.models import optimizers
  from tensor2robot_b200.research.qtopt import pcgrad
  vs = _kat_store()
  opt = pcgrad.PCGrad(optimizers.GradientDescentOptimizer(0.001), allowlist=allowlist, denylist=denylist)
  task_grads = torch.zeros((2, vs.flat.numel()), dtype=torch.float32, device='cuda')
  for t, tg in enumerate(TASK_GRADS):
    for n, g in tg.items():
      v = vs.vars[n]
      task_grads[t, v.offset:v.offset + 2] = torch.from_numpy(g.astype(np.float32)).cuda()
  vs.zero_grad()
  opt.project(vs, task_grads)
  grads = vs.export_tf_grads()
  for i, n in enumerate(NAMES):
    np.testing.assert_allclose(grads[n], KAT_PCGRAD[n] if i in pcgrad_idx else KAT_PLAIN[n], atol=1e-5)
  opt.apply_gradients(vs, 0)
  values = vs.export_tf()
  for i in pcgrad_idx:
    np.testing.assert_allclose(values[NAMES[i]], KAT_RESULT[NAMES[i]], rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize('tasks', [1, 3, 8])
def test_kernel_random_gradients(tasks):
  from tensor2robot_b200 import nn
  from tensor2robot_b200.research.qtopt import pcgrad
  rng = np.random.RandomState(tasks)
  shapes = {'a/w': (3, 3, 64, 64), 'a/b': (64,), 'skip/w': (1000, 37), 'c/tiny': (1,), 'd/big': (300001,)}
  vs = nn.VariableStore('cuda:0', seed=0)
  with nn.variable_store(vs):
    for n, s in shapes.items():
      vs.get_variable(n, s, 0.0)
  vs.finalize()
  opt = pcgrad.PCGrad(None, denylist=['skip/*'])
  host = [{n: rng.standard_normal(s).astype(np.float32) for n, s in shapes.items()} for _ in range(tasks)]
  if tasks > 1:
    host[1]['a/b'][:] = 0            # a task without gradient for one variable
  task_grads = torch.zeros((tasks, vs.flat.numel()), dtype=torch.float32, device='cuda')
  for t in range(tasks):
    for n, g in host[t].items():
      v = vs.vars[n]
      task_grads[t, v.offset:v.offset + v.numel] = torch.from_numpy(g.ravel()).cuda()
  vs.flat_grad.fill_(7.0)            # padding between the variables must stay untouched
  opt.project(vs, task_grads)
  torch.cuda.synchronize()
  want = oracle.compute_gradients(host, list(shapes), denylist=['skip/*'])
  got = vs.export_tf_grads()
  for n in shapes:
    # relative to the task gradients that went in: a 1-element variable with conflicting tasks projects to
    # ~eps * g (a difference of O(1) terms), which fp32 - here as in TF - resolves to a few percent only
    scale = max(np.linalg.norm(want[n]), 1e-2 * sum(np.linalg.norm(h[n]) for h in host))
    err = np.linalg.norm(got[n] - want[n]) / scale
    assert err < 2e-5, (n, err)
  covered = torch.zeros(vs.flat.numel(), dtype=torch.bool)
  for v in vs.trainable_variables():
    covered[v.offset:v.offset + v.numel] = True
  assert bool((vs.flat_grad.cpu()[~covered] == 7.0).all())


@pytest.mark.gpu
def test_compute_gradients_on_a_network():
  """Two task losses (one per pose dimension) on the pose_env regression network: PCGrad.compute_gradients equals the
  oracle applied to the two separately computed task gradients."""
  from tensor2robot_b200 import nn
  from tensor2robot_b200.research.pose_env import pose_env_models as pm
  from tensor2robot_b200.research.qtopt import pcgrad
  from tensor2robot_b200.utils import tensorspec_utils as tu
  rng = np.random.RandomState(0)
  feats = tu.TensorSpecStruct(state=torch.from_numpy(rng.uniform(0, 1, (4, 64, 64, 3)).astype(np.float32)).cuda())
  target = torch.from_numpy(rng.uniform(-1, 1, (4, 2)).astype(np.float32)).cuda()
  model = pm.PoseEnvRegressionModel()
  vs = nn.VariableStore('cuda:0', seed=3)
  with torch.no_grad(), nn.variable_store(vs):
    model.a_func(feats, 'a_func', 'train')
  vs.finalize()

  def losses():
    pose = model.a_func(feats, 'a_func', 'train')['inference_output']
    return [((pose[:, 0] - target[:, 0]) ** 2).mean(), ((pose[:, 1] + 3 * target[:, 1]) ** 2).mean()]

  per_task = []
  for t in range(2):
    with nn.variable_store(vs):
      task_loss = losses()[t]
      vs.zero_grad()
      task_loss.backward()
    per_task.append({k: v.astype(np.float64) for k, v in vs.export_tf_grads().items()})
  opt = pcgrad.PCGrad(None, shuffle=False)
  with nn.variable_store(vs):
    opt.compute_gradients(losses(), vs)
  torch.cuda.synchronize()
  got = vs.export_tf_grads()
  want = oracle.compute_gradients(per_task, list(got))
  conflicts = 0
  for n in got:
    conflicts += float(np.sum(per_task[0][n] * per_task[1][n])) < 0
    err = np.linalg.norm(got[n] - want[n]) / max(np.linalg.norm(want[n]), 1e-30)
    assert err < 1e-4, (n, err)
  assert conflicts > 0, 'the test needs at least one variable with conflicting task gradients'


@pytest.mark.gpu
def test_train_step_with_task_losses(tmp_path):
  """A model whose model_train_fn publishes `pcgrad_losses` trains through the PCGrad wrapper (the reference's
  use_collection_losses route) and moves every variable."""
  from tensor2robot_b200.input_generators import default_input_generator as gens
  from tensor2robot_b200.models import optimizers
  from tensor2robot_b200.research.pose_env import pose_env_models as pm
  from tensor2robot_b200.research.qtopt import pcgrad
  from tensor2robot_b200.utils import train_eval
  calls = []

  class TwoTaskModel(pm.PoseEnvRegressionModel):

    def model_train_fn(self, features, labels, inference_outputs, mode, config=None, params=None):
      pose = inference_outputs['inference_output']
      tasks = [((pose[:, i] - labels.target_pose[:, i]) ** 2).mean() for i in range(2)]
      return tasks[0] + tasks[1], {'pcgrad_losses': tasks}

  class CountingPCGrad(pcgrad.PCGrad):

    def compute_gradients(self, losses, vs):
      calls.append(len(losses))
      return super(CountingPCGrad, self).compute_gradients(losses, vs)

  model = TwoTaskModel(create_optimizer_fn=lambda use_summaries: CountingPCGrad(optimizers.MomentumOptimizer(0.01, 0.9)))
  out = train_eval.train_eval_model(t2r_model=model, input_generator_train=gens.DefaultRandomInputGenerator(batch_size=4),
                                    max_train_steps=2, model_dir=str(tmp_path))
  assert out['global_step'] == 2 and np.isfinite(out['loss']) and calls == [2, 2]
  state = torch.load(str(tmp_path / 'model.ckpt-2.pt'), weights_only=False)
  init = torch.load(str(tmp_path / 'model.ckpt-0.pt'), weights_only=False)
  assert all(np.abs(state['variables'][k] - init['variables'][k]).max() > 0 for k in init['variables'])
