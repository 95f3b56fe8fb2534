"""AbstractPreprocessor: batch preprocessing executed before model_fn
(preprocessors/abstract_preprocessor.py:28-217 of the reference)."""
import abc

from tensor2robot_b200.models import model_interface
from tensor2robot_b200.utils import tensorspec_utils

ModeKeys = model_interface.ModeKeys


class AbstractPreprocessor(abc.ABC):
  """Validates the parsed batch against the in-spec, runs _preprocess_fn, validates the out-spec."""

  def __init__(self, model_feature_specification_fn=None, model_label_specification_fn=None,
               is_model_device_tpu=False):
    for spec_generator in (model_feature_specification_fn, model_label_specification_fn):
      for mode in (ModeKeys.TRAIN, ModeKeys.PREDICT, ModeKeys.EVAL):
        if spec_generator:
          tensorspec_utils.assert_valid_spec_structure(spec_generator(mode))
    self._model_feature_specification_fn = model_feature_specification_fn
    self._model_label_specification_fn = model_label_specification_fn
    self._is_model_device_tpu = is_model_device_tpu

  @property
  def model_feature_specification_fn(self):
    return self._model_feature_specification_fn

  @model_feature_specification_fn.setter
  def model_feature_specification_fn(self, fn):
    self._model_feature_specification_fn = fn

  @property
  def model_label_specification_fn(self):
    return self._model_label_specification_fn

  @model_label_specification_fn.setter
  def model_label_specification_fn(self, fn):
    self._model_label_specification_fn = fn

  @abc.abstractmethod
  def get_in_feature_specification(self, mode):
    """Spec of the features the preprocess_fn consumes."""

  @abc.abstractmethod
  def get_in_label_specification(self, mode):
    """Spec of the labels the preprocess_fn consumes."""

  @abc.abstractmethod
  def get_out_feature_specification(self, mode):
    """Spec of the features the preprocess_fn produces."""

  @abc.abstractmethod
  def get_out_label_specification(self, mode):
    """Spec of the labels the preprocess_fn produces."""

  @abc.abstractmethod
  def _preprocess_fn(self, features, labels, mode):
    """Operates on a BATCH of packed features / labels; may mutate and return them."""

  def preprocess(self, features, labels, mode):
    """Boilerplate packing / validation around _preprocess_fn (abstract_preprocessor.py:171-217).
    Returns flattened (features, labels)."""
    features = tensorspec_utils.validate_and_pack(self.get_in_feature_specification(mode), features,
                                                  ignore_batch=True)
    if labels is not None:
      labels = tensorspec_utils.validate_and_pack(self.get_in_label_specification(mode), labels,
                                                  ignore_batch=True)
    features_preprocessed, labels_preprocessed = self._preprocess_fn(features=features, labels=labels, mode=mode)
    with tensorspec_utils.device_bf16_as_float32():
      features_preprocessed = tensorspec_utils.validate_and_flatten(
          self.get_out_feature_specification(mode), features_preprocessed, ignore_batch=True)
      if labels_preprocessed:
        labels_preprocessed = tensorspec_utils.validate_and_flatten(
            self.get_out_label_specification(mode), labels_preprocessed, ignore_batch=True)
    return features_preprocessed, labels_preprocessed
