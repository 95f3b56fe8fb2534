"""AbstractInputGenerator (input_generators/abstract_input_generator.py:34-160 of the reference)."""
import abc
import functools
import inspect

from tensor2robot_b200.models import model_interface
from tensor2robot_b200.utils import tensorspec_utils

ModeKeys = model_interface.ModeKeys


class AbstractInputGenerator(abc.ABC):
  """Produces batches of (features, labels) fulfilling a model's (preprocessor in-) specs."""

  def __init__(self, batch_size=32):
    self._feature_spec = None
    self._label_spec = None
    self._out_feature_spec = None
    self._out_label_spec = None
    self._preprocess_fn = None
    self._batch_size = batch_size

  @property
  def batch_size(self):
    return self._batch_size

  @batch_size.setter
  def batch_size(self, batch_size):
    self._batch_size = batch_size

  def set_specification_from_model(self, t2r_model, mode):
    """In/out specs and the preprocess function come from the model's preprocessor (:76-98)."""
    preprocessor = t2r_model.preprocessor
    self._feature_spec = preprocessor.get_in_feature_specification(mode)
    tensorspec_utils.assert_valid_spec_structure(self._feature_spec)
    self._label_spec = preprocessor.get_in_label_specification(mode)
    tensorspec_utils.assert_valid_spec_structure(self._label_spec)
    self._out_feature_spec = preprocessor.get_out_feature_specification(mode)
    self._out_label_spec = preprocessor.get_out_label_specification(mode)
    self._preprocess_fn = functools.partial(preprocessor.preprocess, mode=mode)

  def set_feature_specifications(self, feature_spec, out_feature_spec):
    tensorspec_utils.assert_valid_spec_structure(feature_spec)
    self._feature_spec, self._out_feature_spec = feature_spec, out_feature_spec

  def set_label_specifications(self, label_spec, out_label_spec):
    tensorspec_utils.assert_valid_spec_structure(label_spec)
    self._label_spec, self._out_label_spec = label_spec, out_label_spec

  def set_preprocess_fn(self, preprocess_fn):
    if isinstance(preprocess_fn, functools.partial):
      if 'mode' not in preprocess_fn.keywords:
        raise ValueError('The preprocess_fn mode has to be set if a partial function has been passed.')
    elif 'mode' in inspect.getfullargspec(preprocess_fn).args:
      raise ValueError('The passed preprocess_fn has an open argument `mode` which should be patched by a closure or '
                       'with functools.partial.')
    self._preprocess_fn = preprocess_fn

  def _assert_specs_initialized(self):
    if self._feature_spec is None:
      raise ValueError('No feature spec set, please call set_specification_from_model.')
    if self._label_spec is None:
      raise ValueError('No label spec set, please call set_specification_from_model.')

  def create_dataset_input_fn(self, mode):
    """Returns input_fn(params) -> iterator over (features, labels) batches (:131-160)."""
    self._assert_specs_initialized()

    def input_fn(params=None):
      return self._create_dataset(mode=mode, params=params)
    return input_fn

  def create_dataset(self, mode, params=None):
    return self.create_dataset_input_fn(mode)(params)

  @abc.abstractmethod
  def _create_dataset(self, mode, params=None):
    """Iterator over raw (un-preprocessed) host batches keyed by spec path."""
