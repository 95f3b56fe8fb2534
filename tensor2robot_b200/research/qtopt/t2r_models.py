"""QT-Opt T2R models on the B200 engine (research/qtopt/t2r_models.py of the reference).

Class names, constructor arguments, feature/label specifications (keys, names, shapes, dtypes),
output keys and the preprocessor contract follow the reference:
  LegacyGraspingModelWrapper                                   t2r_models.py:62-239
  DefaultGrasping44ImagePreprocessor                           t2r_models.py:242-308
  Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom   t2r_models.py:311-400
plus `ResNet50QCriticModel`, the BASELINE C2/C3 critic (our composition, see resnet_critic.py).
"""
import abc

import torch

from tensor2robot_b200 import nn
from tensor2robot_b200.models import critic_model
from tensor2robot_b200.models import model_interface
from tensor2robot_b200.preprocessors import image_transformations
from tensor2robot_b200.preprocessors import spec_transformation_preprocessor
from tensor2robot_b200.research.qtopt import networks
from tensor2robot_b200.research.qtopt import optimizer_builder
from tensor2robot_b200.research.qtopt import resnet_critic
from tensor2robot_b200.utils import dtypes
from tensor2robot_b200.utils import tensorspec_utils

TRAIN, EVAL, PREDICT = model_interface.TRAIN, model_interface.EVAL, model_interface.PREDICT
INPUT_SHAPE = (512, 640, 3)
TARGET_SHAPE = (472, 472)
TSPEC = tensorspec_utils.ExtendedTensorSpec


def pack_features_kuka_e2e(tf_model, *policy_inputs):
  del tf_model, policy_inputs
  raise NotImplementedError


def log_loss(labels, predictions):
  """tf.losses.log_loss placeholder used as the default `loss_function`: the wrapper computes the
  same quantity from the logits in one fused kernel (see LegacyGraspingModelWrapper.loss_fn)."""
  raise NotImplementedError('log_loss is evaluated by the fused sigmoid+log-loss kernel')


class LegacyGraspingModelWrapper(critic_model.CriticModel, abc.ABC):
  """T2R wrapper around the grasping network definitions."""

  def __init__(self, loss_function=log_loss, learning_rate=1e-4, model_weights_averaging=.9999, momentum=.9,
               export_batch_size=1, use_avg_model_params=True, learning_rate_decay_factor=.999, **kwargs):
    self.hparams = optimizer_builder.HParams(
        batch_size=32, examples_per_epoch=3000000, learning_rate_decay_factor=learning_rate_decay_factor,
        learning_rate=learning_rate, model_weights_averaging=model_weights_averaging, momentum=momentum,
        num_epochs_per_decay=2.0, optimizer='momentum', rmsprop_decay=.9, rmsprop_epsilon=1.0,
        use_avg_model_params=use_avg_model_params)
    self._export_batch_size = export_batch_size
    engine_kwargs = {k: kwargs.pop(k) for k in ('device', 'seed', 'preprocessor_cls') if k in kwargs}
    self.kwargs = kwargs
    super(LegacyGraspingModelWrapper, self).__init__(
        loss_function=loss_function,
        create_optimizer_fn=lambda _: optimizer_builder.BuildOpt(self.hparams),
        action_batch_size=kwargs.get('action_batch_size'),
        use_avg_model_params=use_avg_model_params, **engine_kwargs)
    self._legacy_model = None

  @abc.abstractproperty
  def legacy_model_class(self):
    pass

  def create_legacy_model(self):
    if self._legacy_model is None:
      self._legacy_model = self.legacy_model_class(**self.kwargs)
    return self._legacy_model

  def l2_regularization(self):
    return self.create_legacy_model().l2_regularization

  @abc.abstractmethod
  def pack_features(self, *policy_inputs):
    pass

  def get_trainable_variables(self):
    prefix = self.legacy_model_class.__name__ + '/'
    return [v for v in self.variable_store.trainable_variables() if v.name.startswith(prefix)]

  def get_variables(self):
    prefix = self.legacy_model_class.__name__ + '/'
    return [v for n, v in self.variable_store.vars.items() if n.startswith(prefix)]

  def get_label_specification(self, mode):
    del mode
    return tensorspec_utils.TensorSpecStruct(reward=TSPEC(shape=(1,), dtype=dtypes.float32, name='grasp_success'))

  def get_global_step(self):
    return self.global_step

  def loss_fn(self, features, labels, inference_outputs):
    """tf.losses.log_loss(labels.reward, q_predicted) (t2r_models.py:229-239, critic_model.py:171-192),
    evaluated from the logits by the fused sigmoid+log-loss kernel (same value, eps = 1e-7)."""
    del features
    loss, _ = nn.sigmoid_log_loss(inference_outputs['logits'], labels.reward)
    return loss

  def model_train_fn(self, features, labels, inference_outputs, mode, config=None, params=None):
    """The reference returns tf.losses.get_total_loss() = log loss + l2 regularisation.  The data
    term is returned here for backward(); train_step adds the regularisation term to the reported
    loss and the fused optimizer kernel applies its gradient."""
    del mode, config, params
    return self.loss_fn(features, labels, inference_outputs)


class DefaultGrasping44ImagePreprocessor(spec_transformation_preprocessor.SpecTransformationPreprocessor):
  """In-spec: state/image becomes a (512, 640, 3) uint8 jpeg; TRAIN: random crop + convert +
  photometric distortion, otherwise centre crop + convert (t2r_models.py:242-308)."""

  def _transform_in_feature_specification(self, tensor_spec_struct):
    self.update_spec(tensor_spec_struct, 'state/image', shape=INPUT_SHAPE, dtype=dtypes.uint8, data_format='jpeg')
    return tensor_spec_struct

  def _preprocess_fn(self, features, labels, mode):
    crop = image_transformations.RandomCropImages if mode == TRAIN else image_transformations.CenterCropImages
    image = crop([features.state.image], INPUT_SHAPE, TARGET_SHAPE)[0]
    params = image_transformations.draw_photometric_params() if mode == TRAIN else None
    # convert_image_dtype(float32) + ApplyPhotometricImageDistortions + clip: one kernel, bf16 storage
    out_dtype = torch.float32 if (mode != TRAIN and nn.is_high_precision()) else torch.bfloat16
    features.state.image = image_transformations.convert_and_distort(image, params, out_dtype)
    return features, labels


class Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom(LegacyGraspingModelWrapper):
  """QT-Opt T2R model."""

  def __init__(self, action_batch_size=None, **hparams):
    super(Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom, self).__init__(
        action_batch_size=action_batch_size, **hparams)

  def get_state_specification(self):
    return tensorspec_utils.TensorSpecStruct(image=TSPEC(shape=(472, 472, 3), dtype=dtypes.float32, name='image_1'))

  def get_action_specification(self):
    f32 = dtypes.float32
    return tensorspec_utils.TensorSpecStruct(
        world_vector=TSPEC(shape=(3), dtype=f32, name='world_vector'),
        vertical_rotation=TSPEC(shape=(2), dtype=f32, name='vertical_rotation'),
        close_gripper=TSPEC(shape=(1,), dtype=f32, name='close_gripper'),
        open_gripper=TSPEC(shape=(1,), dtype=f32, name='open_gripper'),
        terminate_episode=TSPEC(shape=(1,), dtype=f32, name='terminate_episode'),
        gripper_closed=TSPEC(shape=(1,), dtype=f32, name='gripper_closed'),
        height_to_bottom=TSPEC(shape=(1,), dtype=f32, name='height_to_bottom'))

  @property
  def default_preprocessor_cls(self):
    return DefaultGrasping44ImagePreprocessor

  def q_func(self, features, scope, mode, config=None, params=None, reuse=True, goal_vector_fn=None,
             goal_spatial_fn=None):
    del scope, config, params, reuse
    base_model = self.create_legacy_model()
    concat_axis = 2 if (mode == PREDICT and self._tile_actions_for_predict) else 1
    images = [None, features.state.image]
    grasp_params = base_model.create_grasp_params_input(features.action.to_dict(), concat_axis)
    logits, end_points = base_model.model(images, grasp_params, goal_spatial_fn=goal_spatial_fn,
                                          goal_vector_fn=goal_vector_fn, is_training=(mode == TRAIN))
    return {'q_predicted': end_points['predictions'], 'logits': logits, 'global_step': self.get_global_step()}

  def pack_features(self, *policy_inputs):
    return pack_features_kuka_e2e(self, *policy_inputs)

  @property
  def legacy_model_class(self):
    return networks.Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom


class ResNet50QCriticModel(Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom):
  """The BASELINE.json C2/C3 critic: same specs / preprocessor / loss, ResNet-50 v2 vision tower."""

  @property
  def legacy_model_class(self):
    return resnet_critic.ResNet50QCritic
