"""CriticModel: Q(state, action) models with continuous actions (models/critic_model.py:43-238)."""
import abc

from tensor2robot_b200.models import abstract_model
from tensor2robot_b200.models import model_interface
from tensor2robot_b200.utils import tensorspec_utils
from tensor2robot_b200.utils import tf_losses

PREDICT = model_interface.PREDICT


class CriticModel(abstract_model.AbstractT2RModel):
  """Critic model with continuous actions trained using MC returns."""

  def __init__(self, loss_function=tf_losses.mean_squared_error, action_batch_size=None, **kwargs):
    super(CriticModel, self).__init__(**kwargs)
    self._loss_function = loss_function
    self._action_batch_size = action_batch_size
    self._tile_actions_for_predict = action_batch_size is not None

  @abc.abstractmethod
  def get_action_specification(self):
    """Specs of the tensors that are unique to each action."""

  @abc.abstractmethod
  def get_state_specification(self):
    """Specs of the tensors shared by all candidate actions."""

  def pack_state_action_to_feature_spec(self, state_params, action_params):
    return tensorspec_utils.TensorSpecStruct(state=state_params, action=action_params)

  def get_feature_specification(self, mode):
    """state + action; in PREDICT mode with action_batch_size the action specs are tiled to
    [action_batch_size] + shape (critic_model.py:106-136)."""
    feature_spec = tensorspec_utils.TensorSpecStruct(state=self.get_state_specification(),
                                                     action=self.get_action_specification())
    if mode == PREDICT and self._tile_actions_for_predict:
      tiled = tensorspec_utils.TensorSpecStruct()
      for key, spec in tensorspec_utils.flatten_spec_structure(self.get_action_specification()).items():
        tiled[key] = tensorspec_utils.ExtendedTensorSpec.from_spec(
            spec, shape=(self._action_batch_size,) + tuple(spec.shape))
      return tensorspec_utils.TensorSpecStruct(state=self.get_state_specification(), action=tiled)
    return feature_spec

  @abc.abstractmethod
  def q_func(self, features, scope, mode, config=None, params=None, reuse=True):
    """Q(state, action).  Returns a {key: tensor} mapping; 'q_predicted' is required."""

  def loss_fn(self, features, labels, inference_outputs):
    """labels.reward vs q_predicted through the configured loss function (critic_model.py:171-192)."""
    del features
    if self._loss_function is None:
      raise ValueError('CriticModel needs a loss_function')
    return self._loss_function(labels=labels.reward, predictions=inference_outputs['q_predicted'])

  def inference_network_fn(self, features, labels, mode, config=None, params=None):
    del labels
    outputs = self.q_func(features=features, mode=mode, scope='q_func', config=config, params=params, reuse=True)
    update_ops = None
    if isinstance(outputs, tuple):
      outputs, update_ops = outputs
    if not isinstance(outputs, dict):
      raise ValueError('The output of q_func is expected to be a dict.')
    if 'q_predicted' not in outputs:
      raise ValueError('For critic models q_predicted is a required key in outputs but is not in {}.'.format(
          list(outputs.keys())))
    return outputs, update_ops

  def model_train_fn(self, features, labels, inference_outputs, mode, config=None, params=None):
    del mode, config, params
    return self.loss_fn(features, labels, inference_outputs)
