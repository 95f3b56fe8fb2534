"""ModelInterface: what infrastructure (input generators, the train/eval driver, predictors) may call
on a T2R model.  Mirrors models/model_interface.py:47-145 of the reference."""
import abc

TRAIN = 'train'
EVAL = 'eval'
PREDICT = 'infer'


class ModeKeys(object):
  """tf.estimator.ModeKeys string values."""
  TRAIN = TRAIN
  EVAL = EVAL
  PREDICT = PREDICT


class ModelInterface(abc.ABC):
  """The model contract: specifications, preprocessor, model_fn."""

  @abc.abstractmethod
  def get_feature_specification_for_packing(self, mode):
    """Feature spec expected by create_pack_features (the preprocessor's in-spec)."""

  @abc.abstractmethod
  def get_label_specification_for_packing(self, mode):
    """Label spec expected by create_pack_features."""

  @abc.abstractproperty
  def preprocessor(self):
    """The preprocessor instance used to convert parsed inputs into model inputs."""

  @abc.abstractmethod
  def get_feature_specification(self, mode):
    """Spec structure (no batch dimension) of the features model_fn consumes."""

  @abc.abstractmethod
  def get_label_specification(self, mode):
    """Spec structure (no batch dimension) of the labels model_fn consumes."""

  @abc.abstractmethod
  def get_run_config(self):
    """Run configuration of the training loop."""

  @abc.abstractmethod
  def model_fn(self, features, labels, mode, config=None, params=None):
    """Runs the model on a batch of packed features / labels."""

  @abc.abstractproperty
  def is_device_tpu(self):
    pass

  @abc.abstractproperty
  def is_device_gpu(self):
    pass

  @abc.abstractproperty
  def device_type(self):
    pass
