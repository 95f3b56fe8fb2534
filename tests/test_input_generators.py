"""input_generators/default_input_generator.py with the cases of the reference's default_input_generator_test.py:33-164:
record / multi-dataset / multi-eval / fractional / weighted / random / constant generators on the real fixture."""
import json
import os

import numpy as np
import pytest

from tensor2robot_b200.input_generators import default_input_generator
from tensor2robot_b200.utils import dtypes
from tensor2robot_b200.utils import tensorspec_utils

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, 'golden', 'pose_env_test_data.tfrecord')
BATCH_SIZE = 2
TSPEC = tensorspec_utils.ExtendedTensorSpec


def _check_input_generator(input_generator):
  feature_spec = tensorspec_utils.TensorSpecStruct()
  feature_spec.state = TSPEC(shape=(64, 64, 3), dtype=dtypes.uint8, name='state/image', data_format='jpeg')
  feature_spec.action = TSPEC(shape=(2,), dtype=dtypes.float32, name='pose')
  label_spec = tensorspec_utils.TensorSpecStruct()
  label_spec.reward = TSPEC(shape=(), dtype=dtypes.float32, name='reward')
  with pytest.raises(ValueError):
    input_generator.create_dataset_input_fn(mode='train')        # no specs yet
  input_generator.set_feature_specifications(feature_spec, feature_spec)
  input_generator.set_label_specifications(label_spec, label_spec)
  np_features, np_labels = next(iter(input_generator.create_dataset_input_fn(mode='train')()))
  np_features = tensorspec_utils.validate_and_pack(feature_spec, np_features, ignore_batch=True)
  np_labels = tensorspec_utils.validate_and_pack(label_spec, np_labels, ignore_batch=True)
  assert list(np_features.state.shape) == [2, 64, 64, 3]
  assert list(np_features.action.shape) == [2, 2]
  assert tuple(np_labels.reward.shape) == (2,)
  return np_features, np_labels


def test_record_input_generator():
  _check_input_generator(default_input_generator.DefaultRecordInputGenerator(file_patterns=FIXTURE, batch_size=BATCH_SIZE))


def test_multi_record_input_generator():
  input_generator = default_input_generator.DefaultRecordInputGenerator(dataset_map={'d1': FIXTURE, 'd2': FIXTURE},
                                                                        batch_size=BATCH_SIZE)
  feature_spec = tensorspec_utils.TensorSpecStruct()
  feature_spec.state = TSPEC(shape=(64, 64, 3), dtype=dtypes.uint8, name='state/image', data_format='jpeg', dataset_key='d1')
  feature_spec.action = TSPEC(shape=(2,), dtype=dtypes.float32, name='pose', dataset_key='d1')
  label_spec = tensorspec_utils.TensorSpecStruct()
  label_spec.reward = TSPEC(shape=(), dtype=dtypes.float32, name='reward', dataset_key='d1')
  label_spec.reward_2 = TSPEC(shape=(), dtype=dtypes.float32, name='reward', dataset_key='d2')
  input_generator.set_feature_specifications(feature_spec, feature_spec)
  input_generator.set_label_specifications(label_spec, label_spec)
  np_features, np_labels = next(iter(input_generator.create_dataset_input_fn(mode='train')()))
  np_features = tensorspec_utils.validate_and_pack(feature_spec, np_features, ignore_batch=True)
  np_labels = tensorspec_utils.validate_and_pack(label_spec, np_labels, ignore_batch=True)
  assert list(np_features.state.shape) == [2, 64, 64, 3] and list(np_features.action.shape) == [2, 2]
  assert tuple(np_labels.reward.shape) == (2,) and tuple(np_labels.reward_2.shape) == (2,)


def test_multi_eval_record_input_generator(monkeypatch):
  monkeypatch.setenv('TF_CONFIG', json.dumps({'multi_eval_name': 'd2'}))
  input_generator = default_input_generator.MultiEvalRecordInputGenerator(eval_map={'d1': FIXTURE, 'd2': 'fubar'},
                                                                          batch_size=2)
  assert input_generator._file_patterns == 'fubar'            # pylint: disable=protected-access
  assert default_input_generator.get_multi_eval_name() == 'd2'


def test_fractional_record_input_generator():
  num_files, fraction = 10, 0.3
  input_generator = default_input_generator.FractionalRecordInputGenerator(
      file_fraction=fraction, file_patterns=','.join([FIXTURE] * num_files), batch_size=BATCH_SIZE)
  assert len(str(input_generator._file_patterns).split(',')) == int(fraction * num_files)   # pylint: disable=protected-access


def test_weighted_record_input_generator():
  _check_input_generator(default_input_generator.WeightedRecordInputGenerator(
      file_patterns=','.join([FIXTURE] * 10), batch_size=BATCH_SIZE))


def test_random_dataset():
  features, _ = _check_input_generator(default_input_generator.DefaultRandomInputGenerator(batch_size=2))
  assert features.state.dtype == np.uint8


def test_constant_dataset():
  features, labels = _check_input_generator(default_input_generator.DefaultConstantInputGenerator(constant_value=1,
                                                                                                  batch_size=2))
  assert (np.asarray(features.action) == 1).all() and (np.asarray(labels.reward) == 1).all()


def test_abstract_input_generator():
  """input_generators/abstract_input_generator_test.py:31-49."""
  import functools
  from tensor2robot_b200.input_generators import abstract_input_generator
  from tensor2robot_b200.preprocessors import noop_preprocessor
  from tensor2robot_b200.utils import mocks
  with pytest.raises(TypeError):
    abstract_input_generator.AbstractInputGenerator()              # pylint: disable=abstract-class-instantiated
  generator = mocks.MockInputGenerator(batch_size=32)
  preprocessor = noop_preprocessor.NoOpPreprocessor()
  with pytest.raises(ValueError):
    generator.set_preprocess_fn(preprocessor.preprocess)           # `mode` is still open
  with pytest.raises(ValueError):
    generator.set_preprocess_fn(functools.partial(preprocessor.preprocess, labels=None))   # a partial without `mode`
  generator.set_preprocess_fn(functools.partial(preprocessor.preprocess, mode='train'))
  generator.set_preprocess_fn(lambda features, labels: (features, labels))
