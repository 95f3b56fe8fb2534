"""SpecTransformationPreprocessor: the in-spec is the model spec with a few entries rewritten
(preprocessors/spec_transformation_preprocessor.py:30-174)."""
from tensor2robot_b200.preprocessors import abstract_preprocessor
from tensor2robot_b200.utils import tensorspec_utils


class SpecTransformationPreprocessor(abstract_preprocessor.AbstractPreprocessor):

  def update_spec(self, tensor_spec_struct, key, **kwargs_for_tensorspec):
    """Rewrites one spec of the (flat) structure in place, e.g. dtype/shape/data_format."""
    tensor_spec_struct[key] = tensorspec_utils.ExtendedTensorSpec.from_spec(tensor_spec_struct[key],
                                                                            **kwargs_for_tensorspec)
    return tensor_spec_struct

  def get_in_feature_specification(self, mode):
    flat = tensorspec_utils.flatten_spec_structure(self._model_feature_specification_fn(mode))
    return self._transform_in_feature_specification(tensorspec_utils.TensorSpecStruct(flat.items()))

  def _transform_in_feature_specification(self, tensor_spec_struct):
    return tensor_spec_struct

  def get_in_label_specification(self, mode):
    flat = tensorspec_utils.flatten_spec_structure(self._model_label_specification_fn(mode))
    return self._transform_in_label_specification(tensorspec_utils.TensorSpecStruct(flat.items()))

  def _transform_in_label_specification(self, tensor_spec_struct):
    return tensor_spec_struct

  def get_out_feature_specification(self, mode):
    return self._model_feature_specification_fn(mode)

  def get_out_label_specification(self, mode):
    return self._model_label_specification_fn(mode)
