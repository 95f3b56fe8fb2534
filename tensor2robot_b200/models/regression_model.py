"""RegressionModel: continuous outputs trained with a mean-squared error (models/regression_model.py:43-200)."""
import abc

from tensor2robot_b200.models import abstract_model
from tensor2robot_b200.utils import tf_losses


class RegressionModel(abstract_model.AbstractT2RModel):
  """Subclasses define `a_func` (and optionally `loss_fn`)."""

  def __init__(self, action_size=2, **kwargs):
    super(RegressionModel, self).__init__(**kwargs)
    self._action_size = action_size

  @abc.abstractmethod
  def a_func(self, features, scope, mode, config=None, params=None, reuse=True):
    """A(state).  Returns a {key: tensor} mapping; 'inference_output' is required."""

  def loss_fn(self, labels, inference_outputs, mode, params=None):
    del mode, params
    return tf_losses.mean_squared_error(labels=labels.target, predictions=inference_outputs['inference_output'])

  def inference_network_fn(self, features, labels, mode, config=None, params=None):
    del labels
    outputs = self.a_func(features=features, mode=mode, scope='a_func', config=config, params=params, reuse=True)
    if not isinstance(outputs, dict):
      raise ValueError('The output of a_func is expected to be a dict.')
    if 'inference_output' not in outputs:
      raise ValueError('For regression models inference_output is a required key in outputs but is not in {}.'.format(
          list(outputs.keys())))
    return outputs

  def model_train_fn(self, features, labels, inference_outputs, mode, config=None, params=None):
    del features, config
    return self.loss_fn(labels, inference_outputs, mode=mode, params=params)

  def create_export_outputs_fn(self, features, inference_outputs, mode, config=None, params=None):
    del features, mode, config, params
    return {'inference_output': inference_outputs['inference_output']}
