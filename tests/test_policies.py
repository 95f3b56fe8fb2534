"""Serving side: policies on a mock predictor (CPU) and CheckpointPredictor + CEMPolicy / RegressionPolicy
on the pose_env models (GPU; predictors/checkpoint_predictor_test.py:41-92)."""
import numpy as np
import pytest

from tensor2robot_b200.policies import policies


class _QuadraticPredictor(object):
  """q(s, a) = -|a - target|^2 and a constant regression output."""

  def __init__(self, target):
    self.target = np.asarray(target, np.float64)
    self.calls = []
    self.global_step = 7

  def predict(self, np_inputs):
    if 'action' in np_inputs:
      a = np.asarray(np_inputs['action'])
      self.calls.append(a.shape)
      return {'q_predicted': -np.sum((a - self.target) ** 2, axis=1)}
    return {'inference_output': self.target[None]}

  def init_randomly(self):
    self.calls.append('init')

  def restore(self):
    self.calls.append('restore')

  model_path = '/some/path'


class _Model(object):

  def pack_features(self, state, context, timestep, actions=None):
    del context, timestep
    out = {'state': np.expand_dims(state, 0)}
    if actions is not None:
      out['action'] = actions
    return out


def test_cem_policy_finds_the_argmax():
  np.random.seed(0)
  predictor = _QuadraticPredictor([0.4, -0.3])
  policy = policies.CEMPolicy(_Model(), action_size=2, cem_iters=5, cem_samples=64, num_elites=10, predictor=predictor)
  action = policy.SelectAction(np.zeros((4, 4, 3)), None, 0)
  assert action.shape == (2,) and np.abs(action - [0.4, -0.3]).max() < 0.05
  assert predictor.calls == [(64, 2)] * 5
  best, debug = policy.get_cem_action(lambda s: -np.sum((s - 1.0) ** 2, axis=1))
  assert set(debug) == {'q_predicted', 'final_params', 'best_idx'} and np.abs(best - 1.0).max() < 0.1
  assert debug['final_params']['stddev'].shape == (2,)
  assert policy.sample_action(np.zeros((4, 4, 3)), 0.5)[1] is None
  policy.init_randomly()
  policy.restore()
  assert predictor.calls[-2:] == ['init', 'restore'] and policy.global_step == 7 and policy.model_path == '/some/path'


def test_regression_policy_and_device_cem_hook():
  predictor = _QuadraticPredictor([0.1, 0.2])
  state = np.zeros((4, 4, 3))
  np.testing.assert_allclose(policies.RegressionPolicy(_Model(), predictor=predictor).SelectAction(state, None, 0), [0.1, 0.2])
  with pytest.raises(NotImplementedError):
    policies.Policy().SelectAction(state, None, 0)
  assert policies.Policy().global_step == 0 and policies.Policy().model_path == 'No model path defined.'
  # device mode: the whole search is delegated (engine.device_cem_selector on a GPU); the predictor is not called
  seen = []
  device = policies.CEMPolicy(_Model(), action_size=2, predictor=predictor,
                              device_maximizer=lambda s: (seen.append(s.shape) or np.array([0.5, -0.5]), 0.9))
  np.testing.assert_allclose(device.SelectAction(state, None, 0), [0.5, -0.5])
  assert seen == [(4, 4, 3)] and predictor.calls == []


def test_predictor_contract_without_gpu():
  from tensor2robot_b200.predictors import checkpoint_predictor
  from tensor2robot_b200.research.pose_env import pose_env_models as pm
  model = pm.PoseEnvRegressionModel()
  with pytest.raises(ValueError):
    checkpoint_predictor.CheckpointPredictor(t2r_model=model, use_gpu=False)
  predictor = checkpoint_predictor.CheckpointPredictor(t2r_model=model)
  with pytest.raises(ValueError):
    predictor.restore()                        # no checkpoint_dir
  with pytest.raises(ValueError):
    predictor.predict({'does_not_matter': np.zeros(1)})
  assert predictor.model_version == -1 and predictor.global_step == -1
  assert list(predictor.get_feature_specification().keys()) == ['state']
  missing = checkpoint_predictor.CheckpointPredictor(t2r_model=model, checkpoint_dir='/random/path/which/does/not/exist',
                                                     timeout=1)
  assert missing.restore() is False


@pytest.mark.gpu
def test_checkpoint_predictor_and_policies_on_pose_env(tmp_path):
  from tensor2robot_b200.input_generators import default_input_generator as gens
  from tensor2robot_b200.predictors import checkpoint_predictor
  from tensor2robot_b200.research.pose_env import pose_env_models as pm
  from tensor2robot_b200.utils import tensorspec_utils
  from tensor2robot_b200.utils import train_eval
  # regression: train 3 steps, restore into a fresh model, same predictions as the trained one
  trained = pm.PoseEnvRegressionModel()
  train_eval.train_eval_model(t2r_model=trained, input_generator_train=gens.DefaultRandomInputGenerator(batch_size=2),
                              max_train_steps=3, model_dir=str(tmp_path))
  model = pm.PoseEnvRegressionModel()
  predictor = checkpoint_predictor.CheckpointPredictor(t2r_model=model, checkpoint_dir=str(tmp_path))
  assert predictor.global_step == -1 and predictor.restore() and predictor.global_step == 3
  assert predictor.model_path.endswith('model.ckpt-3.pt') and predictor.restore()        # unchanged checkpoint
  spec = model.preprocessor.get_in_feature_specification('infer')
  tensorspec_utils.assert_equal(predictor.get_feature_specification(), spec)
  features = tensorspec_utils.make_random_numpy(spec, batch_size=2)
  predictions = predictor.predict(features)
  assert sorted(predictions) == ['inference_output'] and predictions['inference_output'].shape == (2, 2)
  reference = checkpoint_predictor.CheckpointPredictor(t2r_model=trained, checkpoint_dir=str(tmp_path))
  reference.init_randomly()                    # already built and trained: keeps its weights
  np.testing.assert_allclose(reference.predict(features)['inference_output'], predictions['inference_output'], atol=1e-6)
  obs = features['state'][0]
  action = policies.RegressionPolicy(model, predictor=predictor).SelectAction(obs, None, 0)
  np.testing.assert_allclose(action, predictions['inference_output'][0], atol=1e-6)
  predictor.close()
  with pytest.raises(ValueError):
    predictor.predict(features)

  # critic: CEM over Q(image, pose) with randomly initialised weights, 64 samples per iteration through the kernels
  critic = pm.PoseEnvContinuousMCModel()
  q_predictor = checkpoint_predictor.CheckpointPredictor(t2r_model=critic)
  q_predictor.init_randomly()
  seen = []

  def pack_fn(t2r_model, state, context, timestep, samples):
    del t2r_model, context, timestep
    seen.append(samples.shape)
    return {'state/image': np.expand_dims(state, 0), 'action/pose': samples.astype(np.float32)}

  policy = policies.CEMPolicy(critic, action_size=2, cem_iters=3, cem_samples=64, num_elites=10, pack_fn=pack_fn,
                              predictor=q_predictor)
  np.random.seed(0)
  best = policy.SelectAction(obs, None, 0)
  assert best.shape == (2,) and np.isfinite(best).all() and seen == [(64, 2)] * 3
  q = q_predictor.predict(pack_fn(None, obs, None, 0, np.stack([best, best + 0.5])))['q_predicted']
  assert q.shape == (2,) and np.isfinite(q).all()


def test_exported_model_predictor_version_selection(tmp_path):
  """Newest complete numbered export wins; exports in flight (temp dirs, missing assets) are ignored; the specs and the
  global step come from t2r_assets.pbtxt (predictors/exported_savedmodel_predictor.py:52-260)."""
  import os
  from tensor2robot_b200.hooks import td3
  from tensor2robot_b200.predictors import exported_model_predictor as emp
  from tensor2robot_b200.research.pose_env import pose_env_models as pm
  from tensor2robot_b200.utils import dtypes
  from tensor2robot_b200.utils import tensorspec_utils as tu

  class _Store(object):

    def export_tf(self):
      return {'w': np.ones((2, 2), np.float32)}

  class _Model(object):
    variable_store = _Store()

    def __init__(self, step):
      self.global_step = step

    def get_feature_specification_for_packing(self, mode):
      return tu.TensorSpecStruct(state=tu.ExtendedTensorSpec(shape=(64, 64, 3), dtype=dtypes.uint8, name='state/image'))

    def get_label_specification_for_packing(self, mode):
      return tu.TensorSpecStruct(target_pose=tu.ExtendedTensorSpec(shape=(2,), dtype=dtypes.float32, name='target_pose'))

  export_dir = str(tmp_path / 'export')
  assert emp.valid_export_versions(export_dir) == []
  for step in (5, 20, 100):
    td3.export_model(_Model(step), export_dir, step)
  os.makedirs(os.path.join(export_dir, 'temp-0000000200'))
  os.makedirs(os.path.join(export_dir, '0000000300'))                     # started, nothing written yet
  versions = emp.valid_export_versions(export_dir)
  assert [os.path.basename(v) for v in versions] == ['0000000005', '0000000020', '0000000100']

  loaded = []
  predictor = emp.ExportedModelPredictor(export_dir, pm.PoseEnvRegressionModel(), timeout=1)
  assert predictor.global_step == -1 and predictor.model_version == -1
  with pytest.raises(ValueError):
    predictor.get_feature_specification()
  real_load = predictor._load_version                                       # pylint: disable=protected-access

  def spy(path):                    # the weight upload needs a GPU; the assets / bookkeeping part runs here
    loaded.append(os.path.basename(path))
    predictor._ensure_built = lambda: (_ for _ in ()).throw(RuntimeError('stop before the device'))   # pylint: disable=protected-access
    try:
      real_load(path)
    except RuntimeError:
      pass

  predictor._load_version = spy                                             # pylint: disable=protected-access
  assert predictor.restore() and loaded == ['0000000100']
  assert predictor.global_step == 100 and predictor.model_version == 100 and predictor.model_path.endswith('0000000100')
  assert predictor.get_feature_specification().state.shape == (64, 64, 3)
  assert predictor.get_label_specification().target_pose.name == 'target_pose'
  assert predictor.restore() and loaded == ['0000000100']                  # unchanged: not reloaded
  td3.export_model(_Model(400), export_dir, 400)
  assert predictor.restore() and loaded == ['0000000100', '0000000400'] and predictor.global_step == 400
  empty = emp.ExportedModelPredictor(str(tmp_path / 'nothing'), pm.PoseEnvRegressionModel(), timeout=1)
  assert empty.restore() is False
  pinned = emp.ExportedModelPredictor(os.path.join(export_dir, '0000000020'), pm.PoseEnvRegressionModel(), timeout=1)
  assert pinned._newest_version().endswith('0000000020')                    # pylint: disable=protected-access
