"""AbstractPredictor: loads a T2RModel and exposes predict() on numpy features (predictors/abstract_predictor.py:25-81)."""
import abc


class AbstractPredictor(abc.ABC):

  @abc.abstractmethod
  def predict(self, features):
    """features: {key: numpy} -> {key: numpy} model predictions."""

  @abc.abstractmethod
  def get_feature_specification(self):
    """The required input features."""

  def get_label_specification(self):
    return None

  @abc.abstractmethod
  def restore(self):
    """Restores the model parameters from the latest available data."""

  def init_randomly(self):
    """Initialises the model parameters with random values."""

  @abc.abstractmethod
  def close(self):
    """Releases everything held for model evaluation."""

  @abc.abstractmethod
  def assert_is_loaded(self):
    """Raises a ValueError if the predictor has not been restored yet."""

  @property
  def model_version(self):
    return 0

  @property
  def global_step(self):
    return 0

  @property
  def model_path(self):
    return ''
