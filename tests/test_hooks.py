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


def _fake_export_fn(export_dir, global_step):
  path = os.path.join(export_dir, '%010d' % global_step)
  os.makedirs(path, exist_ok=True)
  with open(os.path.join(path, 'weights'), 'w') as f:
    f.write(str(global_step))
  return path


def test_lagged_listener_stays_one_export_behind(tmp_path):
  """hooks/checkpoint_hooks.py:91-201 (the reference's checkpoint_hooks_test scenario): current vs lagged directory
  contents, version garbage collection, and resuming from directories left by an earlier run."""
  from tensor2robot_b200.hooks import checkpoint_hooks
  export_dir, lagged_dir = str(tmp_path / 'export'), str(tmp_path / 'lagged')
  listener = checkpoint_hooks.LaggedCheckpointListener(_fake_export_fn, export_dir, lagged_dir, num_versions=2)
  names = lambda d: sorted(os.listdir(d))
  listener.after_save(None, 10)
  assert names(export_dir) == ['0000000010'] and names(lagged_dir) == ['0000000010']     # nothing older yet
  listener.after_save(None, 20)
  assert names(export_dir) == ['0000000010', '0000000020'] and names(lagged_dir) == ['0000000010']
  listener.after_save(None, 30)
  assert names(export_dir) == ['0000000020', '0000000030']                                # num_versions = 2
  assert names(lagged_dir) == ['0000000010', '0000000020']
  listener.after_save(None, 40)
  assert names(export_dir) == ['0000000030', '0000000040'] and names(lagged_dir) == ['0000000020', '0000000030']
  assert open(os.path.join(lagged_dir, '0000000030', 'weights')).read() == '30'
  # a new listener over the same directories resumes one behind
  resumed = checkpoint_hooks.LaggedCheckpointListener(_fake_export_fn, export_dir, lagged_dir, num_versions=2)
  resumed.after_save(None, 50)
  assert names(export_dir) == ['0000000040', '0000000050'] and names(lagged_dir) == ['0000000030', '0000000040']
  # a lagged directory that fell behind is repaired at construction
  import shutil
  shutil.rmtree(lagged_dir)
  checkpoint_hooks.LaggedCheckpointListener(_fake_export_fn, export_dir, lagged_dir, num_versions=2)
  assert names(lagged_dir) == ['0000000040']
  plain = checkpoint_hooks.CheckpointExportListener(_fake_export_fn, str(tmp_path / 'plain'))
  for step in (1, 2, 3):
    plain.after_save(None, step)
  assert len(names(str(tmp_path / 'plain'))) == 3                                          # no GC without num_versions


def test_td3_hook_builder_triggers_every_save_steps(tmp_path):
  from tensor2robot_b200.hooks import td3
  assert td3.TD3Hooks(export_dir=None, lagged_export_dir=None).create_hooks(None, None) == []
  builder = td3.TD3Hooks(export_dir=str(tmp_path / 'e'), lagged_export_dir=str(tmp_path / 'l'), save_steps=5,
                         num_versions=3, export_fn=_fake_export_fn)
  (hook,) = builder.create_hooks(None, str(tmp_path))
  hook.begin()
  for step in range(1, 18):
    hook.before_step(step - 1)
    hook.after_step(step, 0.0)
  hook.end()
  assert sorted(os.listdir(str(tmp_path / 'e'))) == ['0000000005', '0000000010', '0000000015']
  assert sorted(os.listdir(str(tmp_path / 'l'))) == ['0000000005', '0000000010']


def test_export_model_layout(tmp_path):
  """An export directory: TF-bundle checkpoint (+ `checkpoint` state file) and assets.extra/t2r_assets.pbtxt."""
  from tensor2robot_b200.hooks import td3
  from tensor2robot_b200.utils import dtypes
  from tensor2robot_b200.utils import tensorspec_utils as tu
  from tensor2robot_b200.utils import tf_checkpoint

  class _Store(object):

    def export_tf(self):
      return {'q_func/fc/weights': np.arange(6, dtype=np.float32).reshape(2, 3), 'q_func/fc/biases': np.zeros(3, np.float32)}

  class _Model(object):
    global_step = 12
    variable_store = _Store()

    def get_feature_specification_for_packing(self, mode):
      return tu.TensorSpecStruct(x=tu.ExtendedTensorSpec(shape=(3,), dtype=dtypes.float32, name='measured_position'))

    def get_label_specification_for_packing(self, mode):
      return tu.TensorSpecStruct(y=tu.ExtendedTensorSpec(shape=(1,), dtype=dtypes.float32, name='valid_position'))

  path = td3.export_model(_Model(), str(tmp_path / 'export'), 12)
  assert os.path.basename(path) == '0000000012'
  reader = tf_checkpoint.load_checkpoint(path)
  assert int(reader.get_tensor('global_step')) == 12
  np.testing.assert_array_equal(reader.get_tensor('q_func/fc/weights'), np.arange(6, dtype=np.float32).reshape(2, 3))
  assets = tu.load_t2r_assets_to_file(os.path.join(path, 'assets.extra', 't2r_assets.pbtxt'))
  assert assets.global_step == 12
  spec = tu.TensorSpecStruct.from_proto(assets.feature_spec)
  assert spec.x.shape == (3,) and spec.x.name == 'measured_position'


# ---- the scenarios of the reference's hooks/checkpoint_hooks_test.py:58-178, one by one ---------------------
class _Exporter(object):
  """_MakeSavedModel: export_dir/<id>/savedModel.txt + variables/variables.txt; ids count up per export."""

  def __init__(self, start=0):
    self.export_id = start

  @staticmethod
  def make(export_dir, checkpoint_id):
    path = os.path.join(export_dir, str(checkpoint_id))
    os.makedirs(os.path.join(path, 'variables'))
    for name in ('savedModel.txt', os.path.join('variables', 'variables.txt')):
      with open(os.path.join(path, name), 'w') as f:
        f.write('abc')
    return path

  def __call__(self, export_dir, global_step):
    del global_step
    self.export_id += 1
    return self.make(export_dir, self.export_id)


def _exists(base, checkpoint_id):
  return os.path.exists(os.path.join(base, str(checkpoint_id)))


def test_reference_checkpoint_export_listener_cases(tmp_path):
  from tensor2robot_b200.hooks import checkpoint_hooks
  export_dir = str(tmp_path / 'a')
  listener = checkpoint_hooks.CheckpointExportListener(_Exporter(), export_dir)              # testCheckpointExportListener
  listener.after_save(None, 10)
  assert _exists(export_dir, 1)
  export_dir = str(tmp_path / 'b')
  listener = checkpoint_hooks.CheckpointExportListener(_Exporter(), export_dir, num_versions=3)   # ...GC
  for step in range(5):
    listener.after_save(None, step)
  assert _exists(export_dir, 5) and not _exists(export_dir, 2)
  export_dir = str(tmp_path / 'c')                                                          # ...GCRestore
  os.makedirs(export_dir)
  for step in range(6):
    _Exporter.make(export_dir, step)
  checkpoint_hooks.CheckpointExportListener(_Exporter(), export_dir, num_versions=3)       # the initializer GCs
  assert _exists(export_dir, 5) and not _exists(export_dir, 2)


def _lagged(tmp_path, name, exporter):
  from tensor2robot_b200.hooks import checkpoint_hooks
  export_dir, lagged_dir = str(tmp_path / name / 'export'), str(tmp_path / name / 'lagged_export')
  os.makedirs(export_dir, exist_ok=True)
  os.makedirs(lagged_dir, exist_ok=True)
  make = lambda: checkpoint_hooks.LaggedCheckpointListener(export_fn=exporter, export_dir=export_dir,
                                                           lagged_export_dir=lagged_dir, num_versions=3)
  return export_dir, lagged_dir, make


def test_reference_lagged_listener_cases(tmp_path):
  exporter = _Exporter()                                                                    # testEmptyDir
  export_dir, lagged_dir, make = _lagged(tmp_path, 'empty', exporter)
  listener = make()
  listener.after_save(None, 10)
  assert _exists(lagged_dir, 1) and _exists(export_dir, 1)
  listener.after_save(None, 11)
  assert _exists(export_dir, 2) and not _exists(lagged_dir, 2)          # the lagged policy has not updated
  listener.after_save(None, 12)
  assert _exists(export_dir, 3) and _exists(lagged_dir, 2)              # now it has

  exporter = _Exporter()                                                                    # testInitOneSavedModel
  export_dir, lagged_dir, make = _lagged(tmp_path, 'one', exporter)
  _Exporter.make(export_dir, 1)
  make()
  assert _exists(export_dir, 1) and _exists(lagged_dir, 1)              # the constructor copies the SavedModel over

  exporter = _Exporter(start=1)                                                             # testInitSavedModelUptoDate
  export_dir, lagged_dir, make = _lagged(tmp_path, 'uptodate', exporter)
  _Exporter.make(export_dir, 1)
  _Exporter.make(lagged_dir, 1)
  make().after_save(None, 11)
  assert _exists(export_dir, 2) and not _exists(lagged_dir, 2)

  exporter = _Exporter(start=2)                                                             # ...FromLaggedPosition
  export_dir, lagged_dir, make = _lagged(tmp_path, 'lagged', exporter)
  _Exporter.make(export_dir, 1)
  _Exporter.make(export_dir, 2)
  _Exporter.make(lagged_dir, 1)
  make().after_save(None, 11)
  assert _exists(export_dir, 3) and _exists(lagged_dir, 2)
