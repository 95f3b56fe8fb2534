"""BC-Z image-to-action network (research/bcz/model.py:245-285): FiLM-conditioned ResNet tower + one MLP head
per pose component.  The BCZModel class around it (specs, residual pose assembly, huber / log losses,
research/bcz/model.py:321-950) is not built yet (coverage row A-18)."""
from tensor2robot_b200 import nn
from tensor2robot_b200.layers import bcz_networks
from tensor2robot_b200.layers import resnet

TRAIN = 'train'


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

MIN_GRIPPER_CLOSE = 0.0   # research/bcz/model.py:52-60 (rescaling target: [MIN_GRIPPER_CLOSE, 1])


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


def _weighted(loss, weights):
  """tf.losses.compute_weighted_loss, Reduction.SUM_BY_NONZERO_WEIGHTS: sum(loss * w) / #{w != 0}."""
  w = torch.as_tensor(weights, dtype=loss.dtype, device=loss.device)
  w = torch.broadcast_to(w, loss.shape)
  nonzero = (w != 0).sum().to(loss.dtype)
  return (loss * w).sum() / torch.clamp(nonzero, min=1.0) if nonzero > 0 else (loss * w).sum()


def huber_loss(labels, predictions, weights=1.0, delta=1.0):
  """tf.losses.huber_loss."""
  err = (predictions - labels).abs()
  quad = torch.clamp(err, max=delta)
  return _weighted(0.5 * quad ** 2 + delta * (err - quad), weights)


def mean_squared_error(labels, predictions, weights=1.0):
  return _weighted((predictions - labels) ** 2, weights)


def log_loss(labels, predictions, weights=1.0, epsilon=1e-7):
  """tf.losses.log_loss."""
  return _weighted(-labels * torch.log(predictions + epsilon) - (1 - labels) * torch.log(1 - predictions + epsilon),
                   weights)


def piecewise_scaled_huber(loss_fn, threshold=0.2, slope=0.001):
  def clipped_loss_fn(**kwargs):
    loss = loss_fn(**kwargs)
    return threshold + (loss - threshold) * slope if loss > 1 else loss
  return clipped_loss_fn


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
                     regularization_loss=None):
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
  stop_mask_value = 1.0 - future['stop_token'] if 'stop_token' in future.keys() else 1.0
  train_outputs, nonloss_outputs = {}, {}
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
  if regularization_loss is not None:
    train_outputs['total_regularization_loss'] = regularization_loss
  loss = sum(train_outputs.values())
  train_outputs.update(nonloss_outputs)
  return loss, train_outputs


def xyz_action_trajectory(outputs):
  rotation = outputs['action/quaternion'] if 'action/quaternion' in outputs else outputs['action/axis_angle']
  return torch.cat([outputs['action/xyz'], rotation], dim=-1)
