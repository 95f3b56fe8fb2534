"""Golden-values regression harness (hooks/golden_values_hook_builder.py)."""
import os

import numpy as np
import pytest
import torch


def test_golden_values_hook_records_every_step(tmp_path):
  from tensor2robot_b200.hooks import golden_values_hook_builder as gv
  (hook,) = gv.GoldenValuesHookBuilder().create_hooks(None, str(tmp_path / 'run'))
  hook.begin()
  for step in range(3):
    hook.before_step(step)
    assert gv.get_collection() == {}
    gv.add_golden_tensor(torch.tensor(float(step)), 'xyz_loss')
    gv.add_golden_tensor(torch.tensor([1.0, 2.0]) * step, 'pair')
    hook.after_step(step + 1, None)
  hook.end()
  values = np.load(str(tmp_path / 'run' / 'golden_values.npy'), allow_pickle=True)
  assert len(values) == 3 and sorted(values[2]) == ['pair', 'xyz_loss']
  assert float(values[2]['xyz_loss']) == 2.0 and values[1]['pair'].tolist() == [1.0, 2.0]


def test_bcz_losses_join_the_golden_collection():
  from tensor2robot_b200.hooks import golden_values_hook_builder as gv
  from tensor2robot_b200.research.bcz import model as bcz
  from tensor2robot_b200.research.bcz import pose_components_lib
  from tensor2robot_b200.utils import tensorspec_utils as tu
  gv.clear_collection()
  rng = np.random.RandomState(0)
  comps = pose_components_lib.DEFAULT_ACTION_COMPONENTS
  outputs = {'xyz_residual': torch.from_numpy(rng.standard_normal((2, 3, 3)).astype(np.float32)),
             'quaternion': torch.from_numpy(rng.standard_normal((2, 3, 4)).astype(np.float32)),
             'target_close': torch.sigmoid(torch.from_numpy(rng.standard_normal((2, 3, 1)).astype(np.float32))),
             'quaternion_norm': torch.ones(2, 3, 1)}
  labels = tu.TensorSpecStruct()
  labels['future/xyz_residual'] = torch.zeros(2, 3, 3)
  labels['future/quaternion'] = torch.zeros(2, 3, 4)
  labels['future/target_close'] = torch.ones(2, 3, 1)
  loss, train_outputs = bcz.training_outputs(labels, outputs, comps)
  collection = gv.get_collection()
  assert set(collection) == set(train_outputs) and 'xyz_loss' in collection
  assert float(loss) == pytest.approx(sum(float(v) for k, v in train_outputs.items() if k.endswith('_loss')), rel=1e-6)


def test_train_eval_rejects_foreign_hooks():
  from tensor2robot_b200.research.pose_env import pose_env_models as pm
  from tensor2robot_b200.utils import train_eval
  with pytest.raises(NotImplementedError):
    train_eval.train_eval_model(t2r_model=pm.PoseEnvRegressionModel(), train_hook_builders=[object()])
  with pytest.raises(NotImplementedError):
    train_eval.train_eval_model(t2r_model=pm.PoseEnvRegressionModel(), create_exporters_fn=lambda *a: [])


@pytest.mark.gpu
def test_golden_values_from_a_training_run(tmp_path):
  from tensor2robot_b200.hooks import golden_values_hook_builder as gv
  from tensor2robot_b200.input_generators import default_input_generator as gens
  from tensor2robot_b200.research.bcz import model as bcz
  from tensor2robot_b200.utils import train_eval
  pre = lambda **kw: bcz.BCZPreprocessor(image_size=(96, 96), crop_size=(104, 128), input_size=(112, 144), **kw)
  model = bcz.BCZModel(image_size=(96, 96), input_size=(112, 144), resnet_size=18, num_waypoints=2, preprocessor_cls=pre)
  out = train_eval.train_eval_model(t2r_model=model, input_generator_train=gens.DefaultRandomInputGenerator(batch_size=4),
                                    max_train_steps=2, model_dir=str(tmp_path),
                                    train_hook_builders=[gv.GoldenValuesHookBuilder()])
  values = np.load(os.path.join(str(tmp_path), 'golden_values.npy'), allow_pickle=True)
  assert len(values) == 2 and 'xyz_loss' in values[0] and 'quaternion_norm_loss' in values[0]
  total = sum(float(v) for k, v in values[1].items() if k.endswith('_loss'))
  assert np.isfinite(total) and abs(total - out['loss']) < 1e-4 * max(1.0, abs(total))
