"""BC-Z (research/bcz/model.py): FiLM-conditioned ResNet image-to-action network with one MLP head per pose
component (:245-285), pose assembly (:321-460), weighted huber / log losses (:476-585), the BCZPreprocessor
(:69-196, mixup included) and the BCZModel T2R class (:641-950).  Not built: cutout (the reference raises too), eval metrics."""
from tensor2robot_b200 import nn
from tensor2robot_b200.layers import bcz_networks
from tensor2robot_b200.layers import resnet

TRAIN = 'train'


def spatial_softmax_network(features, is_training, pose_components, num_waypoints, condition_input=None):
  """Spatial-softmax image-to-action network (model.py:196-242): the vision_layers tower, the task embedding
  concatenated to the feature points, one pose MLP emitting every component of every waypoint.  Returns
  ({component: fp32 [B, num_waypoints, size]}, feature_points)."""
  from tensor2robot_b200.layers import vision_layers
  image = features.image if hasattr(features, 'image') else features['image']
  with nn.variable_scope('vision_model'):
    feature_points, _ = vision_layers.BuildImagesToFeaturesModel(nn.to_f32(image), is_training=is_training)
    if condition_input is not None:
      feature_points = torch.cat([feature_points, condition_input.to(feature_points.dtype)], -1)
    action_sizes = [t[1] for t in pose_components]
    estimated_pose, _ = vision_layers.BuildImageFeaturesToPoseModel(
        feature_points, aux_input=None, aux_output_dim=0, num_outputs=sum(action_sizes) * num_waypoints)
  network_output_dict = {}
  i = 0
  for name, size, is_residual, _ in pose_components:
    if is_residual:
      name += '_residual'
    n = size * num_waypoints
    network_output_dict[name] = estimated_pose[..., i:i + n].reshape(-1, num_waypoints, size)
    i += n
  return network_output_dict, feature_points


def resnet_film_network(features, mode, pose_components, num_waypoints, film_generator_fn=None,
                        condition_input=None, concat_cond_image=None, fc_layers=(100, 100), resnet_size=50):
  """features.image: bf16 [B, h, w, 3] preprocessed frames; condition_input: fp32 [B, E] task embedding fed to
  `film_generator_fn` (e.g. layers.resnet.linear_film_generator).  Returns ({component name: fp32
  [B, num_waypoints, size], 'policy_image_features': [B, F]}, state_features [B, C3])."""
  if concat_cond_image is not None:
    raise NotImplementedError('conditioning images concatenated on the channel axis are not built')
  is_training = mode == TRAIN
  image = features.image if hasattr(features, 'image') else features['image']
  with nn.variable_scope('vision_model'):
    outputs = resnet.resnet_model(image, is_training, num_classes=1, resnet_size=resnet_size,
                                  return_intermediate_values=True, film_generator_fn=film_generator_fn,
                                  film_generator_input=condition_input)
    net = nn.to_f32(outputs['final_reduce_mean'])
    action_sizes, names = [], []
    for name, size, is_residual, _ in pose_components:
      names.append(name + '_residual' if is_residual else name)
      action_sizes.append(size)
    estimated_components = bcz_networks.MultiHeadMLP(net, action_sizes, num_waypoints, fc_layers, is_training)
    # block_layer3 is used to optionally infer the task
    state_features = nn.to_f32(nn.global_mean(outputs['block_layer3']))
    network_output_dict = dict(zip(names, estimated_components))
    network_output_dict['policy_image_features'] = net
    return network_output_dict, state_features


# ---------------------------------------------------------------------------------------------
# Pose assembly and losses (research/bcz/model.py:321-585).  Small [B, waypoints, k] tensors: plain torch
# on whatever device the network outputs live on (host-scale logic, like the spec utilities).
# ---------------------------------------------------------------------------------------------
import torch  # pylint: disable=wrong-import-position

MIN_GRIPPER_CLOSE = 0.2                       # research/bcz/model.py:58-60
GRIPPER_CLOSE_FRACTION_TO_OPEN_GRIPPER = 0.4
NUM_DEBUG_TASKS = 21


def quaternion_multiply(q1, q2):
  """tensorflow_graphics.geometry.transformation.quaternion.multiply, [x, y, z, w] layout (absent third-party
  dependency; the Hamilton product restated)."""
  x1, y1, z1, w1 = q1.unbind(-1)
  x2, y2, z2, w2 = q2.unbind(-1)
  x = x1 * w2 + y1 * z2 - z1 * y2 + w1 * x2
  y = -x1 * z2 + y1 * w2 + z1 * x2 + w1 * y2
  z = x1 * y2 - y1 * x2 + z1 * w2 + w1 * z2
  w = -x1 * x2 - y1 * y2 - z1 * z2 + w1 * w2
  return torch.stack([x, y, z, w], -1)


from tensor2robot_b200.hooks import golden_values_hook_builder  # pylint: disable=wrong-import-position
from tensor2robot_b200.utils import tf_losses  # pylint: disable=wrong-import-position

_weighted = tf_losses.compute_weighted_loss
huber_loss = tf_losses.huber_loss
mean_squared_error = tf_losses.mean_squared_error
log_loss = tf_losses.log_loss


def piecewise_scaled_huber(loss_fn, threshold=0.2, slope=0.001):
  def clipped_loss_fn(**kwargs):
    loss = loss_fn(**kwargs)
    return threshold + (loss - threshold) * slope if loss > 1 else loss
  return clipped_loss_fn


def predict_stop_network(state_embedding, fc_layers=(100, 100), num_waypoints=1, scope_name='predict_stop'):
  """Small MLP predicting (continue, fail / help, success) logits [B, num_waypoints * 3] from the state embedding
  (model.py:286-317): slim.stack of fully_connected + layer_norm + ReLU, a 3-way head, and - behind a stop_gradient -
  the heads of the remaining waypoints."""
  with nn.variable_scope(scope_name):
    net = nn.to_f32(state_embedding)
    for i, units in enumerate(fc_layers):
      scope = 'Stack/fully_connected_%d' % (i + 1)
      net = nn.dense_f32(net, units, scope=scope, bias_rows=0)        # slim drops the bias under a normaliser
      net = nn.layer_norm(net, scope=scope + '/LayerNorm', relu=True)
    logits = nn.dense_f32(net, 3, scope='fully_connected')
    if num_waypoints > 1:
      rest_logits = nn.dense_f32(net.detach(), (num_waypoints - 1) * 3, scope='fully_connected_1')
      logits = torch.cat([logits, rest_logits], dim=-1)
  return logits


def compute_stop_state_loss(stop_state_labels, stop_state_predictions, class_weights):
  """tf.losses.softmax_cross_entropy of the one-hot stop-state labels, each example weighted by its class weight
  (model.py:462-473; `class_weights` is gin.REQUIRED in the reference)."""
  weights = (stop_state_labels * torch.as_tensor(class_weights, dtype=stop_state_labels.dtype,
                                                 device=stop_state_labels.device)).sum(-1)
  ce = -(stop_state_labels * torch.log_softmax(stop_state_predictions.float(), dim=-1)).sum(-1)
  return tf_losses.compute_weighted_loss(ce, weights)


def infer_outputs(features, network_output_dict, action_components, rescale_target_close):
  """Network head outputs -> absolute pose components (model.py:321-460): residual components are added to
  the present pose, quaternions are normalised (and composed with the present one when residual), gripper /
  stop logits go through a sigmoid.  `network_output_dict['quaternion']` is overwritten by the normalised
  quaternion like in the reference."""
  inference_outputs = {}
  action_outputs = []
  present = features.present if hasattr(features, 'present') else features['present']
  for name, size, is_residual, _ in action_components:
    predict_name = name + '_residual' if is_residual else name
    value = network_output_dict[predict_name]
    batch_dims = list(value.shape[:-2])

    def current(n, k):
      return present[n].reshape(batch_dims + [1, k]).to(value.dtype)

    if name == 'quaternion':
      quaternion_norm = torch.linalg.norm(value, dim=-1, keepdim=True)
      quaternion = value / quaternion_norm
      if is_residual:
        quaternion = quaternion_multiply(current(name, 4), quaternion)
      action_outputs.append(quaternion)
      network_output_dict['quaternion'] = quaternion
      inference_outputs['quaternion_norm'] = quaternion_norm
    elif name in ('target_close', 'stop_token'):
      if is_residual:
        raise ValueError('target_close/stop_token do not support residual gripper')
      value = torch.sigmoid(value)
      if rescale_target_close:
        value = MIN_GRIPPER_CLOSE + value * (1 - MIN_GRIPPER_CLOSE)
      action_outputs.append(value)
    elif name == 'base_joystick_xy':
      action_outputs.append(torch.tanh(value))
    elif name == 'arm_joints_velocity':
      action_outputs.append(value)
    elif name in ('xyz', 'axis_angle', 'arm_joints', 'pantilt', 'robot_linear_velocity', 'robot_angular_velocity'):
      action_outputs.append(value + current(name, size) if is_residual else value)
    else:
      raise ValueError('unknown action component %r' % name)
  inference_outputs.update(network_output_dict)
  for (name, _, _, _), output in zip(action_components, action_outputs):
    inference_outputs['action/' + name] = output
  inference_outputs['action_trajectory'] = torch.cat(action_outputs, dim=-1)
  for key in ('image', 'depth_image'):
    if key in features.keys():
      inference_outputs[key] = features[key]
  return inference_outputs


def training_outputs(labels, network_output_dict, action_components, quaternion_penalty=0.01, loss_name='huber',
                     regularization_loss=None, stop_state_class_weights=None):
  """Per-component regression / log losses with the component weights, masked after the stop token, plus the
  QuaterNet norm penalty (model.py:476-585).  Returns (loss, train_outputs)."""
  if loss_name == 'mse':
    reg_loss_fn = mean_squared_error
  elif loss_name == 'huber':
    reg_loss_fn = huber_loss
  elif loss_name == 'clipped_huber':
    reg_loss_fn = lambda **kw: torch.clamp(huber_loss(**kw), 0.0, 6.0)
  elif loss_name == 'piecewise_scaled_huber':
    reg_loss_fn = piecewise_scaled_huber(loss_fn=huber_loss)
  else:
    raise ValueError('invalid loss')
  future = labels.future if hasattr(labels, 'future') else labels['future']
  stop_token = future['stop_token'] if 'stop_token' in future.keys() else None
  train_outputs, nonloss_outputs = {}, {}
  fused_kind = {'huber': 'huber', 'mse': 'mse'}.get(loss_name)
  n_segments = 2 * len(action_components) + (1 if 'quaternion_norm' in network_output_dict else 0)
  if (fused_kind is not None and n_segments <= 16 and
      all(network_output_dict[n + '_residual' if r else n].is_cuda for n, _, r, _ in action_components)):
    # every component loss, its first-waypoint diagnostic and the norm penalty: ONE launch (csrc/losses.cu)
    mask = stop_token.reshape(-1) if stop_token is not None else None
    specs, names = [], []
    for name, _, is_residual, weight in action_components:
      key = name + '_residual' if is_residual else name
      predicted = network_output_dict[key]
      kind = 'sigmoid_log' if name in ('target_close', 'stop_token') else fused_kind
      waypoints = predicted.shape[-2]
      specs.append(dict(kind=kind, predictions=predicted, labels=future[key], weight=weight, row_mask=mask,
                        complement=True))
      names.append(name + '_loss')
      specs.append(dict(kind=kind, predictions=predicted, labels=future[key], weight=weight, row_mod=waypoints,
                        in_total=False, differentiable=False))
      names.append('first_' + name + '_error')
    if 'quaternion_norm' in network_output_dict:
      specs.append(dict(kind=fused_kind, predictions=network_output_dict['quaternion_norm'], labels=1.0,
                        weight=quaternion_penalty, row_mask=mask, complement=True))
      names.append('quaternion_norm_loss')
    losses, sigmoids = nn.weighted_losses(specs)
    fused_total, fused_names = losses[len(specs)], set(n for n, sp in zip(names, specs) if sp.get('in_total', True))
    sigmoids = iter(sigmoids)
    for i, (spec, out_name) in enumerate(zip(specs, names)):
      (train_outputs if spec.get('in_total', True) else nonloss_outputs)[out_name] = losses[i]
      if spec['kind'] == 'sigmoid_log':
        q = next(sigmoids)
        if spec.get('in_total', True):
          nonloss_outputs[out_name[:-len('_loss')] + '_predicted'] = q
  else:
    fused_total, fused_names = None, set()
    stop_mask_value = 1.0 - stop_token if stop_token is not None else 1.0
    for name, _, is_residual, weight in action_components:
      key = name + '_residual' if is_residual else name
      predicted = network_output_dict[key]
      label = future[key].to(predicted.dtype)
      if name in ('target_close', 'stop_token'):
        predicted = torch.sigmoid(predicted)
        nonloss_outputs[name + '_predicted'] = predicted
        loss_fn = log_loss
      else:
        loss_fn = reg_loss_fn
      stop_mask = stop_mask_value * torch.ones_like(predicted)
      train_outputs[name + '_loss'] = loss_fn(labels=label, predictions=predicted, weights=weight * stop_mask)
      nonloss_outputs['first_' + name + '_error'] = loss_fn(labels=label[..., 0, :], predictions=predicted[..., 0, :],
                                                            weights=weight)
    if 'quaternion_norm' in network_output_dict:
      predicted = network_output_dict['quaternion_norm']
      train_outputs['quaternion_norm_loss'] = reg_loss_fn(labels=torch.ones_like(predicted), predictions=predicted,
                                                          weights=quaternion_penalty * stop_mask_value)
  if 'stop_state' in network_output_dict:          # stop state prediction loss (model.py:566-573)
    if stop_state_class_weights is None:
      raise ValueError('compute_stop_state_loss.class_weights is required (gin.REQUIRED in the reference)')
    stop_labels = _one_hot(labels.future.stop_state.reshape(-1), 3)       # tf.one_hot: out-of-range -> zero row
    train_outputs['stop_state_loss'] = compute_stop_state_loss(stop_labels, network_output_dict['stop_state'],
                                                               stop_state_class_weights)
  if regularization_loss is not None:
    train_outputs['total_regularization_loss'] = regularization_loss
  loss = sum(v for k, v in train_outputs.items() if k not in fused_names)   # the kernel already summed its terms
  if fused_total is not None:
    loss = fused_total + loss
  train_outputs.update(nonloss_outputs)
  for name, tensor in train_outputs.items():          # each of the losses joins the golden collection (:581-583)
    golden_values_hook_builder.add_golden_tensor(tensor, name)
  return loss, train_outputs


def xyz_action_trajectory(outputs):
  rotation = outputs['action/quaternion'] if 'action/quaternion' in outputs else outputs['action/axis_angle']
  return torch.cat([outputs['action/xyz'], rotation], dim=-1)



# ---------------------------------------------------------------------------------------------
# Preprocessor and T2R model (research/bcz/model.py:63-196, 641-950)
# ---------------------------------------------------------------------------------------------
import enum  # pylint: disable=wrong-import-position
import inspect  # pylint: disable=wrong-import-position

import numpy as np  # pylint: disable=wrong-import-position

from tensor2robot_b200.layers import resnet as resnet_layers  # pylint: disable=wrong-import-position
from tensor2robot_b200.models import abstract_model  # pylint: disable=wrong-import-position
from tensor2robot_b200.preprocessors import distortion  # pylint: disable=wrong-import-position
from tensor2robot_b200.preprocessors import spec_transformation_preprocessor  # pylint: disable=wrong-import-position
from tensor2robot_b200.research.bcz import pose_components_lib  # pylint: disable=wrong-import-position
from tensor2robot_b200.utils import dtypes  # pylint: disable=wrong-import-position
from tensor2robot_b200.utils import tensorspec_utils  # pylint: disable=wrong-import-position

EVAL, PREDICT = 'eval', 'infer'
TensorSpec = tensorspec_utils.ExtendedTensorSpec
_RNG = np.random.RandomState(0)
_NOISE_GENERATORS = {}     # device -> torch.Generator of the task-embedding noise, seeded from _RNG on first use


def _noise_generator(device):
  key = (str(device), id(_RNG))
  gen = _NOISE_GENERATORS.get(key)
  if gen is None:
    gen = _NOISE_GENERATORS[key] = torch.Generator(device=device)
    gen.manual_seed(int(_RNG.randint(0, 2**31 - 1)))
  return gen


def _one_hot(ids, depth):
  """tf.one_hot: ids outside [0, depth) give an all-zero row."""
  return (ids.long()[:, None] == torch.arange(depth, device=ids.device)[None, :]).float()


class ConditionMode(enum.Enum):
  ONEHOT_TASKID = 1
  LANGUAGE_EMBEDDING = 2


def mixup_reverse(x, lmbda):
  """lmbda * x + (1 - lmbda) * tf.reverse(x, axis=[0]) on a [B, ...] tensor (fp32 on the device kernel)."""
  import ctypes as C
  from tensor2robot_b200 import _lib
  if not x.is_cuda:
    return lmbda * x + (1.0 - lmbda) * torch.flip(x, dims=[0])
  xf = x.float().contiguous()
  y = torch.empty_like(xf)
  _lib.call('t2r_mixup_reverse_f32', C.c_void_p(xf.data_ptr()), C.c_void_p(y.data_ptr()), xf.shape[0],
            xf.numel() // xf.shape[0], float(lmbda), _lib.current_stream_ptr())
  return y.to(x.dtype) if x.dtype.is_floating_point else y


class BCZPreprocessor(spec_transformation_preprocessor.SpecTransformationPreprocessor):
  """Image conversion / crop / resize for single frames (model.py:69-196)."""

  def __init__(self, image_size=(100, 100), crop_size=(512, 640), input_size=(512, 640), is_sequence=False,
               mixup_alpha=0.0, cutout_size=0, mock_subtask=False, binarize_gripper=True, rescale_gripper=False,
               image_distortion_fn=None, **kwargs):
    self._mixup_alpha = float(mixup_alpha)
    self._image_size = tuple(image_size)
    self._crop_size = tuple(crop_size)
    self._input_size = tuple(input_size)
    self._is_sequence = is_sequence
    self._cutout_size = cutout_size
    self._mock_subtask = mock_subtask
    self._binarize_gripper = binarize_gripper
    self._rescale_gripper = rescale_gripper
    self._image_distortion_fn = image_distortion_fn
    super(BCZPreprocessor, self).__init__(**kwargs)

  @property
  def rescale_gripper(self):
    return self._rescale_gripper

  def get_in_feature_specification(self, mode):
    flat = tensorspec_utils.flatten_spec_structure(self._model_feature_specification_fn(mode))
    flat = tensorspec_utils.TensorSpecStruct(flat.items())
    for key in ('original_image', 'original_depth_image'):   # produced by _preprocess_fn, never parsed
      if mode != PREDICT and key in flat.keys():
        del flat[key]
    return self._transform_in_feature_specification(flat)

  def _transform_in_feature_specification(self, flat_spec_structure):
    self.update_spec(flat_spec_structure, 'image', shape=self._input_size + (3,), dtype=dtypes.uint8,
                     data_format='jpeg')
    return flat_spec_structure

  def _preprocess_fn(self, features, labels, mode):
    features.original_image = features.image
    image = distortion.preprocess_image(features.image, mode, self._is_sequence, input_size=self._input_size,
                                        target_size=self._image_size, crop_size=self._crop_size,
                                        image_distortion_fn=self._image_distortion_fn)
    if self._mixup_alpha > 0.0 and labels is not None and mode == TRAIN:
      # Mixup (model.py:164-172): ONE lambda ~ Beta(alpha, alpha) per batch blends every sample with the batch
      # reversed, image and future labels alike
      lmbda = float(_RNG.beta(self._mixup_alpha, self._mixup_alpha))
      image = mixup_reverse(image, lmbda)
      for key in list(labels.future.keys()):
        labels.future[key] = mixup_reverse(labels.future[key], lmbda)
    features.image = nn.to_bf16(image)          # the tower computes in bf16
    if self._cutout_size > 0 and mode == TRAIN:
      raise NotImplementedError('Open-source BC-Z Model does not support cutout augmentation.')
    key = 'target_close'
    if labels is not None and key in labels.future.keys():
      if self._binarize_gripper:
        labels.future[key] = (labels.future[key] > GRIPPER_CLOSE_FRACTION_TO_OPEN_GRIPPER).to(labels.future[key].dtype)
      if self._rescale_gripper:
        labels.future[key] = torch.clamp((labels.future[key] - MIN_GRIPPER_CLOSE) / (1 - MIN_GRIPPER_CLOSE), min=0.)
    if self._mock_subtask:
      features.subtask_id = torch.zeros_like(features.subtask_id)
    return features, labels


def get_gripper_accuracy_metrics(inference_outputs, features, labels):
  """Closing / opening prediction of the first waypoint against the sensed gripper state
  (research/bcz/model.py:588-617): accuracy, AUC, precision, recall and the positive frequency of both events."""
  from tensor2robot_b200.utils import metrics
  key = 'target_close'
  current = features.present[key].float()
  predicted = inference_outputs[key][:, 0].float() - current
  label = labels.future[key][:, 0].float() - current
  out = {}
  for name, lab, pred in (('closing', label > 0, predicted > 0), ('opening', label < 0, predicted < 0)):
    lab, pred = lab.float(), pred.float()
    out[name + '_accuracy'] = metrics.accuracy(lab, pred)
    out[name + '_auc'] = metrics.auc(lab, pred)
    out[name + '_precision'] = metrics.precision(lab, pred)
    out[name + '_recall'] = metrics.recall(lab, pred)
    out[name + '_pos_freq'] = metrics.accuracy(torch.ones_like(lab), lab)
  return out


class BCZModel(abstract_model.AbstractT2RModel):
  """Single-image configurable regression model for BC-Z (model.py:641-950)."""

  def __init__(self, state_components=None, action_components=None, predict_stop=False, image_size=(100, 100),
               input_size=None, dataset_keys=None, num_waypoints=1, num_past=0, num_total_users=0,
               network_fn=resnet_film_network, ignore_task_embedding=False, task_embedding_noise_std=0.1,
               init_checkpoint=None, mask_stop_token=False, cond_modality=ConditionMode.ONEHOT_TASKID,
               film_generator_fn=resnet_layers.linear_film_generator, resnet_size=50, stop_state_class_weights=None,
               **kwargs):
    super(BCZModel, self).__init__(**kwargs)
    self._predict_stop = predict_stop
    self._stop_state_class_weights = stop_state_class_weights
    self._image_size = tuple(image_size)
    self._input_size = tuple(input_size) if input_size else None
    self._dataset_keys = dataset_keys
    self._num_waypoints = num_waypoints
    self._num_past = num_past
    self._network_fn = network_fn
    self._ignore_task_embedding = ignore_task_embedding
    self._task_embedding_noise_std = task_embedding_noise_std
    self._action_components = action_components or pose_components_lib.DEFAULT_ACTION_COMPONENTS
    self._state_components = state_components or []
    self._init_checkpoint = init_checkpoint
    self._mask_stop_token = mask_stop_token
    self._num_total_users = num_total_users
    self._cond_mode = cond_modality
    self._film_generator_fn = film_generator_fn      # gin: resnet_film_network.film_generator_fn
    self._resnet_size = resnet_size

  @property
  def default_preprocessor_cls(self):
    return BCZPreprocessor

  @property
  def action_component_names(self):
    return [p[0] for p in self._action_components]

  @property
  def is_joint_space(self):
    return 'arm_joints' in self.action_component_names

  @property
  def is_xyz_space(self):
    return 'xyz' in self.action_component_names

  def pack_features(self, state, prev_episode_data, timestep):
    del prev_episode_data, timestep
    return state

  def get_feature_specification(self, mode):
    del mode
    f32 = dtypes.float32
    features = tensorspec_utils.TensorSpecStruct()
    features.image = TensorSpec(shape=self._image_size + (3,), dtype=f32, name='present/image/encoded',
                                data_format='jpeg', is_sequence=False)
    present = tensorspec_utils.TensorSpecStruct()
    for name, size, _ in self._state_components:
      present[name] = TensorSpec(shape=(size,), dtype=f32, name='present/' + name, is_sequence=False)
    for name, size, _, _ in self._action_components:
      data_name = 'sensed_close' if name == 'target_close' else name   # target_close holds future information
      present[name] = TensorSpec(shape=(size,), dtype=f32, name='present/' + data_name, is_sequence=False)
    features.present = present
    if self._cond_mode == ConditionMode.ONEHOT_TASKID:
      features.subtask_id = TensorSpec(shape=(1,), dtype=dtypes.int64, name='subtask_id')
    elif self._cond_mode == ConditionMode.LANGUAGE_EMBEDDING:
      features.sentence_embedding = TensorSpec(shape=(512,), dtype=f32, name='sentence_embedding')
    if self._num_total_users:
      features.user_id = TensorSpec(shape=(1,), dtype=dtypes.int64, name='user_int')
    features.camera_intrinsics = TensorSpec(shape=(3, 3), dtype=f32, name='present/camera_rgb/intrinsics',
                                            is_optional=True)
    features.camera_pose_base = TensorSpec(shape=(12,), dtype=f32, name='present/camera_pose_base', is_optional=True)
    input_size = self._input_size if self._input_size else (512, 640)
    features.original_image = TensorSpec(shape=input_size + (3,), dtype=dtypes.uint8, data_format='jpeg',
                                         is_optional=True)
    if self._num_past:
      past = tensorspec_utils.TensorSpecStruct()
      for name, size, residual in self._state_components:
        past[name + '_residual' if residual else name] = TensorSpec(
            shape=(self._num_past, size), dtype=f32, name='past/' + (name + '_residual' if residual else name),
            is_sequence=False)
      features.past = past
    return features

  def get_label_specification(self, mode):
    del mode
    future = tensorspec_utils.TensorSpecStruct()
    for name, size, residual, _ in self._action_components:
      key = name + '_residual' if residual else name
      future[key] = TensorSpec(shape=(self._num_waypoints, size), dtype=dtypes.float32, name='future/' + key,
                               is_sequence=False)
    if self._predict_stop:
      future['stop_state'] = TensorSpec(shape=(), dtype=dtypes.int64, name='present/stop_state')
    if self._mask_stop_token:
      future.stop_token = TensorSpec(shape=(self._num_waypoints, 1), dtype=dtypes.float32, name='future/stop_token',
                                     is_sequence=False)
    return tensorspec_utils.TensorSpecStruct(future=future)

  def augment_condition_input(self, condition_input, features, is_training):
    if self._task_embedding_noise_std is not None and is_training:
      # tf.random.normal on the embedding (model.py:815-817), drawn on the device: a host draw of [B, 512] normals plus its
      # pageable H2D copy cost ~2 ms of a 13.5 ms step and synchronised the launching thread
      if condition_input.is_cuda:
        noise = torch.randn(condition_input.shape, dtype=torch.float32, device=condition_input.device,
                            generator=_noise_generator(condition_input.device))
      else:
        noise = torch.from_numpy(_RNG.standard_normal(tuple(condition_input.shape)).astype(np.float32))
      condition_input = condition_input + noise * self._task_embedding_noise_std
    if self._ignore_task_embedding:
      condition_input = None
    extra = []
    if self._state_components:
      extra.append(torch.cat([features.present[t[0]].float() for t in self._state_components], dim=-1))
    if self._num_total_users:
      extra.append(_one_hot(features.user_id[:, 0], self._num_total_users))
    if self._num_past:
      prev = torch.cat([features.past[n + '_residual' if r else n].float() for n, _, r in self._state_components], -1)
      extra.append(prev.reshape(prev.shape[0], -1))
    for e in extra:
      condition_input = e if condition_input is None else torch.cat([condition_input, e.to(condition_input.device)], -1)
    return condition_input

  def inference_network_fn(self, features, labels, mode, config=None, params=None):
    del config, params
    is_training = mode == TRAIN
    if self._cond_mode == ConditionMode.ONEHOT_TASKID:
      condition_input = _one_hot(features.subtask_id[:, 0], NUM_DEBUG_TASKS)
    else:
      condition_input = features.sentence_embedding.float()
    condition_input = self.augment_condition_input(condition_input, features, is_training)
    # what gin binds in the reference (resnet_film_network.film_generator_fn / resnet_size) are constructor arguments
    accepted = inspect.signature(self._network_fn).parameters
    extra = {}
    if 'film_generator_fn' in accepted:
      extra['film_generator_fn'] = self._film_generator_fn if condition_input is not None else None
    if 'resnet_size' in accepted:
      extra['resnet_size'] = self._resnet_size
    network_outputs_dict, state_embedding = self._network_fn(
        features, mode, self._action_components, self._num_waypoints, condition_input=condition_input, **extra)
    outputs = infer_outputs(features, network_outputs_dict, self._action_components,
                            self.preprocessor.rescale_gripper)
    if self._predict_stop:
      outputs['stop_state'] = predict_stop_network(state_embedding)
    if not self._ignore_task_embedding:
      outputs['condition_input'] = condition_input
    return outputs

  def model_train_fn(self, features, labels, inference_outputs, mode, config=None, params=None):
    del features, mode, config, params
    return training_outputs(labels, inference_outputs, self._action_components,
                            stop_state_class_weights=self._stop_state_class_weights)

  def model_eval_fn(self, features, labels, inference_outputs, train_loss, train_outputs, mode, config=None,
                    params=None):
    """Streaming means of every train output, the stop-state accuracy and the gripper open / close classification
    metrics (research/bcz/model.py:894-929)."""
    del train_loss, mode, config, params
    from tensor2robot_b200.utils import metrics
    out = {}
    if train_outputs is not None:
      for key, value in train_outputs.items():
        out['mean_' + key] = metrics.mean(value)
    if self._predict_stop:
      predictions = torch.argmax(inference_outputs['stop_state'], dim=-1)
      out['accuracy_stop_state'] = metrics.accuracy(labels.future.stop_state.long(), predictions)
    if train_outputs and labels is not None and 'target_close' in self.action_component_names:
      out.update(get_gripper_accuracy_metrics(inference_outputs, features, labels))
    return out
