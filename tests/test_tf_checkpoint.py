"""TensorFlow tensor-bundle checkpoints against the reference's own fixture
test_data/mock_exported_savedmodel/variables (copied verbatim to tests/golden/mock_savedmodel_variables): every block
and tensor checksum written by TensorFlow verifies, the writer reproduces the data shard and every BundleEntryProto
byte for byte, and a model initialises from a checkpoint by variable name (models/abstract_model.py:87-126)."""
import os
import shutil

import numpy as np
import pytest

from tensor2robot_b200.utils import tf_checkpoint as tc

HERE = os.path.dirname(os.path.abspath(__file__))
PREFIX = os.path.join(HERE, 'golden', 'mock_savedmodel_variables', 'variables')


def test_reads_the_reference_fixture():
  reader = tc.load_checkpoint(PREFIX)
  shapes = reader.get_variable_to_shape_map()
  assert len(shapes) == 21 and shapes['global_step'] == [] and shapes['MockT2RModel.dense.0/kernel'] == [3, 32]
  assert shapes['MockT2RModel.batch_norm.1/moving_variance'] == [16] and shapes['MockT2RModel.dense.4/kernel'] == [8, 1]
  dtypes = reader.get_variable_to_dtype_map()
  assert dtypes['global_step'] == 'int64' and dtypes['MockT2RModel.dense.0/bias'] == 'float32'
  assert reader.has_tensor('MockT2RModel.dense.2/bias') and not reader.has_tensor('MockT2RModel.dense.3/bias')
  assert int(reader.get_tensor('global_step')) == 1100
  np.testing.assert_array_equal(reader.get_tensor('MockT2RModel.batch_norm.0/moving_variance'), np.ones(32, np.float32))
  gamma = reader.get_tensor('MockT2RModel.batch_norm.0/gamma')
  assert gamma.dtype == np.float32 and abs(float(gamma.mean()) - 1.0) < 0.05      # trained a little away from 1
  with pytest.raises(KeyError):
    reader.get_tensor('nope')
  assert tc.load_checkpoint(os.path.dirname(PREFIX)).has_tensor('global_step')        # directory form


def test_checksums_catch_corruption(tmp_path):
  for suffix in ('.index', '.data-00000-of-00001'):
    shutil.copy(PREFIX + suffix, str(tmp_path / ('variables' + suffix)))
  prefix = str(tmp_path / 'variables')
  with open(prefix + '.data-00000-of-00001', 'r+b') as f:
    f.seek(100)
    b = f.read(1)
    f.seek(100)
    f.write(bytes([b[0] ^ 1]))
  reader = tc.CheckpointReader(prefix)
  bad = 0
  for name in reader.get_variable_to_shape_map():
    try:
      reader.get_tensor(name)
    except tc.CheckpointError:
      bad += 1
  assert bad == 1
  with open(prefix + '.index', 'r+b') as f:
    f.seek(20)
    b = f.read(1)
    f.seek(20)
    f.write(bytes([b[0] ^ 4]))
  with pytest.raises(tc.CheckpointError):
    tc.CheckpointReader(prefix)
  with open(prefix + '.index', 'wb') as f:
    f.write(b'not a table' * 10)
  with pytest.raises(tc.CheckpointError):
    tc.CheckpointReader(prefix)


def test_snappy_decoder():
  # literal only; literal + overlapping 1-byte-offset copy (run-length); 2-byte-offset copy; long literal (60 tag)
  assert tc.snappy_decompress(b'\x05\x10hello') == b'hello'
  assert tc.snappy_decompress(b'\x09\x00a' + bytes([((8 - 4) << 2) | 1, 1])) == b'a' * 9
  assert tc.snappy_decompress(b'\x0a\x10abcde' + bytes([((5 - 1) << 2) | 2, 5, 0])) == b'abcdeabcde'
  body = bytes(range(256)) * 2
  assert tc.snappy_decompress(b'\x80\x04' + bytes([61 << 2]) + (511).to_bytes(2, 'little') + body) == body
  with pytest.raises(tc.CheckpointError):
    tc.snappy_decompress(b'\x05\x00a' + bytes([1, 9]))          # back-reference before the start
  with pytest.raises(tc.CheckpointError):
    tc.snappy_decompress(b'\x06\x10hello')                      # length mismatch


def test_writer_reproduces_the_fixture_encoding(tmp_path):
  reader = tc.load_checkpoint(PREFIX)
  tensors = {n: reader.get_tensor(n) for n in reader.get_variable_to_shape_map()}
  out = str(tmp_path / 'ckpt' / 'model.ckpt-1100')
  tc.write_checkpoint(out, tensors)
  with open(PREFIX + '.data-00000-of-00001', 'rb') as f, open(out + '.data-00000-of-00001', 'rb') as g:
    assert f.read() == g.read()                                  # same tensor order and bytes as BundleWriter
  assert tc.raw_index_entries(out) == tc.raw_index_entries(PREFIX)   # header + every BundleEntryProto byte for byte
  again = tc.load_checkpoint(out)
  for n, a in tensors.items():
    np.testing.assert_array_equal(again.get_tensor(n), a)
    assert again.get_tensor(n).dtype == a.dtype


def test_large_bundle_round_trip(tmp_path):
  """Many variables -> several index data blocks; odd dtypes and shapes."""
  rng = np.random.RandomState(0)
  tensors = {'scope_%03d/w' % i: rng.standard_normal((3, i % 7 + 1)).astype(np.float32) for i in range(400)}
  tensors['step'] = np.array(-5, np.int64)
  tensors['flags'] = np.array([True, False, True])
  tensors['half'] = rng.standard_normal((2, 2)).astype(np.float16)
  tensors['empty'] = np.zeros((0, 4), np.float32)
  prefix = str(tmp_path / 'big')
  tc.write_checkpoint(prefix, tensors)
  assert os.path.getsize(prefix + '.index') > 3 * 4096
  reader = tc.load_checkpoint(prefix)
  assert sorted(reader.get_variable_to_shape_map()) == sorted(tensors)
  for n, a in tensors.items():
    got = reader.get_tensor(n)
    assert got.dtype == a.dtype and got.shape == a.shape
    np.testing.assert_array_equal(got, a)


@pytest.mark.gpu
def test_model_round_trip_through_a_tf_checkpoint(tmp_path):
  """Train 2 steps, export a TF bundle, initialise a fresh model from it by variable name: same predictions."""
  import torch
  from tensor2robot_b200.input_generators import default_input_generator as gens
  from tensor2robot_b200.models import abstract_model
  from tensor2robot_b200.predictors import checkpoint_predictor
  from tensor2robot_b200.research.pose_env import pose_env_models as pm
  from tensor2robot_b200.utils import tensorspec_utils
  from tensor2robot_b200.utils import train_eval
  trained = pm.PoseEnvRegressionModel()
  train_eval.train_eval_model(t2r_model=trained, input_generator_train=gens.DefaultRandomInputGenerator(batch_size=2),
                              max_train_steps=2, model_dir=str(tmp_path / 'run'))
  prefix = train_eval.save_tf_checkpoint(trained, str(tmp_path / 'tf'))
  reader = tc.load_checkpoint(str(tmp_path / 'tf'))                    # via the `checkpoint` state file
  assert int(reader.get_tensor('global_step')) == 2 and prefix.endswith('model.ckpt-2')
  assert reader.get_variable_to_shape_map()['a_func/state_features/conv2/weights'] == [3, 3, 3, 32]
  fresh = pm.PoseEnvRegressionModel(init_from_checkpoint_fn=abstract_model.default_init_from_checkpoint_fn(prefix), seed=5)
  features = tensorspec_utils.make_random_numpy(fresh.preprocessor.get_in_feature_specification('infer'), batch_size=3)
  p_fresh = checkpoint_predictor.CheckpointPredictor(t2r_model=fresh)
  p_fresh.init_randomly()                                              # builds, then the init fn overwrites the values
  p_trained = checkpoint_predictor.CheckpointPredictor(t2r_model=trained)
  p_trained.init_randomly()
  np.testing.assert_allclose(p_fresh.predict(features)['inference_output'], p_trained.predict(features)['inference_output'],
                             atol=1e-6)
  # a variable missing in the checkpoint: error unless partial restores are allowed; filters skip variables
  partial = {n: reader.get_tensor(n) for n in reader.get_variable_to_shape_map() if 'pose_fc2' not in n}
  tc.write_checkpoint(str(tmp_path / 'partial' / 'ckpt'), partial)
  broken = pm.PoseEnvRegressionModel(
      init_from_checkpoint_fn=abstract_model.default_init_from_checkpoint_fn(str(tmp_path / 'partial' / 'ckpt')))
  with pytest.raises(ValueError):
    checkpoint_predictor.CheckpointPredictor(t2r_model=broken).init_randomly()
  seen = []
  tolerant = pm.PoseEnvRegressionModel(init_from_checkpoint_fn=abstract_model.default_init_from_checkpoint_fn(
      str(tmp_path / 'partial' / 'ckpt'), allow_partial_restore=True,
      filter_restorables_fn=lambda v: seen.append(v.op.name) or 'LayerNorm' not in v.name))
  checkpoint_predictor.CheckpointPredictor(t2r_model=tolerant).init_randomly()
  assert 'a_func/pose_fc2/weights' in seen
  values = tolerant.variable_store.export_tf()
  np.testing.assert_array_equal(values['a_func/state_features/conv2/weights'], reader.get_tensor('a_func/state_features/conv2/weights'))
  np.testing.assert_array_equal(values['a_func/pose_fc0/LayerNorm/gamma'], np.ones(100, np.float32))    # filtered out
  del torch


def test_mutated_indexes_only_raise_checkpoint_errors(tmp_path):
  """Fuzz: byte flips, truncations and insertions in the index (TensorFlow's snappy blocks and this module's
  uncompressed ones), with checksum verification off so that the mutations reach the decoders."""
  reader = tc.load_checkpoint(PREFIX)
  tc.write_checkpoint(str(tmp_path / 'w'), {n: reader.get_tensor(n) for n in reader.get_variable_to_shape_map()})
  seeds = [open(PREFIX + '.index', 'rb').read(), open(str(tmp_path / 'w.index'), 'rb').read()]
  shutil.copy(PREFIX + '.data-00000-of-00001', str(tmp_path / 'f.data-00000-of-00001'))
  rng = np.random.RandomState(0)
  parsed = rejected = 0
  for it in range(1500):
    b = bytearray(seeds[it % 2])
    mode = rng.randint(3)
    if mode == 0:
      for _ in range(rng.randint(1, 5)):
        b[rng.randint(len(b))] = rng.randint(256)
    elif mode == 1:
      b = b[:rng.randint(1, len(b))]
    else:
      i = rng.randint(len(b))
      b[i:i] = bytes(rng.randint(0, 256, rng.randint(1, 9)).astype(np.uint8))
    with open(str(tmp_path / 'f.index'), 'wb') as f:
      f.write(bytes(b))
    try:
      mutated = tc.CheckpointReader(str(tmp_path / 'f'), verify=False)
      for name in list(mutated.get_variable_to_shape_map())[:30]:
        try:
          mutated.get_tensor(name)
        except (tc.CheckpointError, KeyError):
          pass
      parsed += 1
    except tc.CheckpointError:
      rejected += 1
  assert parsed > 50 and rejected > 50
