"""NoOpPreprocessor: in and out specs are the model's specs (preprocessors/noop_preprocessor.py)."""
from tensor2robot_b200.preprocessors import abstract_preprocessor
from tensor2robot_b200.utils import tensorspec_utils


class NoOpPreprocessor(abstract_preprocessor.AbstractPreprocessor):

  def get_in_feature_specification(self, mode):
    return tensorspec_utils.flatten_spec_structure(self._model_feature_specification_fn(mode))

  def get_in_label_specification(self, mode):
    return tensorspec_utils.flatten_spec_structure(self._model_label_specification_fn(mode))

  def get_out_feature_specification(self, mode):
    return tensorspec_utils.flatten_spec_structure(self._model_feature_specification_fn(mode))

  def get_out_label_specification(self, mode):
    return tensorspec_utils.flatten_spec_structure(self._model_label_specification_fn(mode))

  def _preprocess_fn(self, features, labels, mode):
    return features, labels
