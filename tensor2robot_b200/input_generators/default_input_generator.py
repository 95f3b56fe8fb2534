"""Default input generators (input_generators/default_input_generator.py:48-314 of the reference).

Batches are produced on the host as numpy arrays keyed by spec path; the training driver stages them
to the device (pinned memory + async H2D on a copy stream) and runs the model's preprocessor there.
"""
import json
import os

import numpy as np

from tensor2robot_b200.input_generators import abstract_input_generator
from tensor2robot_b200.models import model_interface
from tensor2robot_b200.utils import tensorspec_utils
from tensor2robot_b200.utils import tfdata

ModeKeys = model_interface.ModeKeys


def _get_tf_config_env():
  return json.loads(os.environ.get('TF_CONFIG', '{}'))


def get_multi_eval_name(tf_config_env=None):
  tf_config_env = tf_config_env or _get_tf_config_env()
  return tf_config_env.get('multi_eval_name')


class DefaultRecordInputGenerator(abstract_input_generator.AbstractInputGenerator):
  """Reads TFRecord files of serialized tf.Examples (:48-101)."""

  def __init__(self, file_patterns=None, dataset_map=None, label='', seed=None, shard=(0, 1), **parent_kwargs):
    super(DefaultRecordInputGenerator, self).__init__(**parent_kwargs)
    if file_patterns and dataset_map:
      raise ValueError('Only one of `file_patterns` or `dataset_map` should be set.')
    self._file_patterns = file_patterns
    self._dataset_map = dataset_map
    self._label = label
    self._seed, self._shard = seed, shard

  def _create_dataset(self, mode, params=None):
    if not (self._file_patterns or self._dataset_map):
      raise ValueError('The file patterns nor dataset_map are set. File patterns: {} Dataset map: {}'.format(
          self._file_patterns, self._dataset_map))
    return tfdata.default_input_fn_tmpl(
        file_patterns=self._file_patterns or self._dataset_map,
        batch_size=tfdata.get_batch_size(params, self._batch_size), feature_spec=self._feature_spec,
        label_spec=self._label_spec, mode=mode, seed=self._seed, shard=self._shard)


class FractionalRecordInputGenerator(DefaultRecordInputGenerator):
  """Uses only the first `file_fraction` of the files (:105-125)."""

  def __init__(self, file_fraction=1.0, **parent_kwargs):
    super(FractionalRecordInputGenerator, self).__init__(**parent_kwargs)
    if file_fraction < 1.0:
      _, filenames = tfdata.get_data_format_and_filenames(self._file_patterns)
      n = int(file_fraction * len(filenames))
      self._file_patterns = 'tfrecord:' + ','.join(filenames[:n])


class MultiEvalRecordInputGenerator(DefaultRecordInputGenerator):
  """Picks the eval dataset named by TF_CONFIG's multi_eval_name (:128-140)."""

  def __init__(self, eval_map=None, **parent_kwargs):
    super(MultiEvalRecordInputGenerator, self).__init__(**parent_kwargs)
    multi_eval_name = get_multi_eval_name()
    if eval_map and multi_eval_name:
      self._file_patterns = eval_map[multi_eval_name]


class GeneratorInputGenerator(abstract_input_generator.AbstractInputGenerator):
  """Batches from a python generator of single examples (:143-193)."""

  def __init__(self, sequence_length=None, **kwargs):
    self._sequence_length = sequence_length
    super(GeneratorInputGenerator, self).__init__(**kwargs)

  def _generator_fn(self, batch_size):
    raise NotImplementedError

  def _create_dataset(self, mode, params=None):
    del mode
    batch_size = tfdata.get_batch_size(params, self._batch_size)
    for features, labels in self._generator_fn(batch_size):
      yield features, labels


class DefaultRandomInputGenerator(GeneratorInputGenerator):
  """Random batches shaped by the specs: uniform[0,255] integers, [0,1) floats (:197-206)."""

  def _generator_fn(self, batch_size):
    while True:
      yield (tensorspec_utils.make_random_numpy(self._feature_spec, batch_size, self._sequence_length),
             tensorspec_utils.make_random_numpy(self._label_spec, batch_size, self._sequence_length))


class DefaultConstantInputGenerator(GeneratorInputGenerator):
  """Constant batches (:210-226)."""

  def __init__(self, constant_value, **kwargs):
    self._constant_value = constant_value
    super(DefaultConstantInputGenerator, self).__init__(**kwargs)

  def _generator_fn(self, batch_size):
    while True:
      yield (tensorspec_utils.make_constant_numpy(self._feature_spec, self._constant_value, batch_size,
                                                  self._sequence_length),
             tensorspec_utils.make_constant_numpy(self._label_spec, self._constant_value, batch_size,
                                                  self._sequence_length))


class WeightedRecordInputGenerator(DefaultRecordInputGenerator):
  """Samples each batch's examples from several file patterns with given weights (:229-314)."""

  def __init__(self, file_patterns, weights=None, seed=None, **parent_kwargs):
    super(WeightedRecordInputGenerator, self).__init__(file_patterns=file_patterns, seed=seed, **parent_kwargs)
    self._weights = weights

  def _create_dataset(self, mode, params=None):
    batch_size = tfdata.get_batch_size(params, self._batch_size)
    _, filenames_list = tfdata.get_data_format_and_filenames_list(self._file_patterns)
    weights = self._weights or [1.0 / len(filenames_list)] * len(filenames_list)
    if len(weights) != len(filenames_list):
      raise ValueError('Weights need to be same length as number of filenames.')
    weights = np.asarray(weights, np.float64) / np.sum(weights)
    rng = np.random.RandomState(self._seed)
    streams = [tfdata.shuffled(tfdata.record_stream(f, mode, self._seed, self._shard), tfdata.SHUFFLE_BUFFER_SIZE,
                               self._seed) for f in filenames_list]
    parse_fn = tfdata.create_parse_tf_example_fn(self._feature_spec, self._label_spec)
    while True:
      choice = rng.choice(len(streams), size=batch_size, p=weights)   # sample_from_datasets(weights, seed)
      try:
        records = [next(streams[c]) for c in choice]
      except StopIteration:
        return
      yield parse_fn(records)
