"""T2R models of the pose toy environment (research/pose_env/pose_env_models.py:40-325): the Monte-Carlo critic
Q(image, pose) and the image -> pose regression model, with their uint8 -> float32 preprocessors.  Both networks are
32-channel fp32 layers on 64x64 frames (csrc/vision_small.cu); slim's arg_scope defaults
(research/dql_grasping_lib/tf_modules.py:25-44: LayerNorm, ReLU, truncated_normal(0.01), stride 2, VALID) are spelled
out per layer, and variables keep the reference's auto-generated scope names (Conv, Conv_1, Stack/fully_connected_1)."""
import numpy as np
import torch

from tensor2robot_b200 import nn
from tensor2robot_b200.layers import vision_layers
from tensor2robot_b200.models import critic_model
from tensor2robot_b200.models import regression_model
from tensor2robot_b200.preprocessors import abstract_preprocessor
from tensor2robot_b200.preprocessors import image_transformations
from tensor2robot_b200.utils import dtypes
from tensor2robot_b200.utils import tensorspec_utils
from tensor2robot_b200.utils import tf_losses

TensorSpec = tensorspec_utils.ExtendedTensorSpec
TRAIN, PREDICT = 'train', 'infer'


def _to_float(image):
  """tf.image.convert_image_dtype(uint8 -> float32) on the device (a no-op on floats)."""
  if image.dtype == torch.uint8:
    return image_transformations.convert_and_distort(image, None, torch.float32)
  return image.float()


class _ModelSpecPreprocessor(abstract_preprocessor.AbstractPreprocessor):
  """Label and output specs are the model's own (pose_env_models.py:62-79)."""

  def get_in_label_specification(self, mode):
    return tensorspec_utils.flatten_spec_structure(self._model_label_specification_fn(mode))

  def get_out_feature_specification(self, mode):
    return tensorspec_utils.flatten_spec_structure(self._model_feature_specification_fn(mode))

  def get_out_label_specification(self, mode):
    return tensorspec_utils.flatten_spec_structure(self._model_label_specification_fn(mode))


class DefaultPoseEnvContinuousPreprocessor(_ModelSpecPreprocessor):
  """Converts the state image from uint8 to float32 (pose_env_models.py:40-89)."""

  def get_in_feature_specification(self, mode):
    model_spec = self._model_feature_specification_fn(mode)
    feature_spec = tensorspec_utils.TensorSpecStruct()
    image = model_spec.state.image
    feature_spec['state/image'] = TensorSpec(shape=image.shape, dtype=dtypes.uint8, name=image.name,
                                             data_format=image.data_format)
    feature_spec['action/pose'] = model_spec.action.pose
    return feature_spec

  def _preprocess_fn(self, features, labels, mode):
    features.state.image = _to_float(features.state.image)
    return features, labels


class PoseEnvContinuousMCModel(critic_model.CriticModel):
  """Continuous MC critic for the pose env (pose_env_models.py:92-181)."""

  def get_action_specification(self):
    return tensorspec_utils.TensorSpecStruct(pose=TensorSpec(shape=(2,), dtype=dtypes.float32, name='pose'))

  def get_state_specification(self):
    return tensorspec_utils.TensorSpecStruct(
        image=TensorSpec(shape=(64, 64, 3), dtype=dtypes.float32, name='state/image', data_format='jpeg'))

  @property
  def default_preprocessor_cls(self):
    return DefaultPoseEnvContinuousPreprocessor

  def get_label_specification(self, mode):
    del mode
    return tensorspec_utils.TensorSpecStruct(reward=TensorSpec(shape=(), dtype=dtypes.float32, name='reward'))

  def _q_features(self, state, action, is_training=True, reuse=True):
    """[B,64,64,3] x [Bc,2] -> [Bc, h*w*32]: three 3x3 / 2 VALID conv + LayerNorm + ReLU layers, the action embedded
    by a plain FC (ReLU, bias: outside the arg_scope) and added at every position of the tiled feature map."""
    del is_training, reuse
    net = state
    channels = 32
    init = nn.truncated_normal(0.01)
    with nn.variable_scope('q_features'):
      for layer_index in range(3):
        scope = 'Conv' if layer_index == 0 else 'Conv_{}'.format(layer_index)
        net = nn.conv2d_f32(net, channels, 3, stride=2, padding='VALID', use_bias=False, scope=scope, initializer=init)
        net = nn.layer_norm(net, scope=scope + '/LayerNorm', relu=True)
      action_context = nn.relu(nn.dense_f32(action, channels, scope='fully_connected', regularize=False))
      net = nn.tile_add_context(net, action_context)
      net = net.reshape(net.shape[0], -1)
    return net

  def q_func(self, features, scope, mode, config=None, params=None, reuse=True):
    del config, params
    is_training = mode == TRAIN
    with nn.variable_scope(scope):
      image = _to_float(features.state.image)
      pose = features.action.pose.float()
      pose = pose.reshape(-1, pose.shape[-1])      # PREDICT with tiled actions: [B, A, 2] -> [B*A, 2]
      net = self._q_features(image, pose, is_training=is_training, reuse=reuse)
      for i in (1, 2):
        net = nn.relu(nn.dense_f32(net, 100, scope='Stack/fully_connected_{}'.format(i), regularize=False))
      net = nn.dense_f32(net, 1, scope='fully_connected', regularize=False)
      return {'q_predicted': net.squeeze(1)}

  def pack_features(self, state, context, timestep, actions):
    del context, timestep
    return tensorspec_utils.TensorSpecStruct(state=np.expand_dims(state, 0), action=actions)


class DefaultPoseEnvRegressionPreprocessor(_ModelSpecPreprocessor):
  """Converts the state image from uint8 to float32 (pose_env_models.py:184-228)."""

  def get_in_feature_specification(self, mode):
    state = self._model_feature_specification_fn(mode).state
    feature_spec = tensorspec_utils.TensorSpecStruct()
    feature_spec['state'] = TensorSpec(shape=state.shape, dtype=dtypes.uint8, name=state.name,
                                       data_format=state.data_format)
    return feature_spec

  def _preprocess_fn(self, features, labels, mode):
    features.state = _to_float(features.state)
    return features, labels


class PoseEnvRegressionModel(regression_model.RegressionModel):
  """Continuous regression output model for the pose env (pose_env_models.py:231-325)."""

  @property
  def default_preprocessor_cls(self):
    return DefaultPoseEnvRegressionPreprocessor

  def get_feature_specification(self, mode):
    del mode
    return tensorspec_utils.TensorSpecStruct(
        state=TensorSpec(shape=(64, 64, 3), dtype=dtypes.float32, name='state/image', data_format='jpeg'))

  def get_label_specification(self, mode):
    del mode
    return tensorspec_utils.TensorSpecStruct(
        target_pose=TensorSpec(shape=(self._action_size,), dtype=dtypes.float32, name='target_pose'),
        reward=TensorSpec(shape=(1,), dtype=dtypes.float32, name='reward'))

  def pack_features(self, state, context, timestep):
    del context, timestep
    return tensorspec_utils.TensorSpecStruct(state=np.expand_dims(state, 0))

  @property
  def action_size(self):
    return self._action_size

  def a_func(self, features, scope, mode, config=None, params=None, reuse=True, context_fn=None):
    del config, params, reuse
    is_training = mode == TRAIN
    image = _to_float(features.state)
    with nn.variable_scope(scope):
      with nn.variable_scope('state_features'):
        feature_points, _ = vision_layers.BuildImagesToFeaturesModel(image, is_training=is_training)
      if context_fn:
        feature_points = context_fn(feature_points)
      estimated_pose, _ = vision_layers.BuildImageFeaturesToPoseModel(feature_points, num_outputs=self._action_size)
    return {'inference_output': estimated_pose, 'state_features': feature_points}

  def loss_fn(self, labels, inference_outputs, mode, params=None):
    del mode, params
    return tf_losses.mean_squared_error(labels=labels.target_pose, predictions=inference_outputs['inference_output'],
                                        weights=labels.reward)
