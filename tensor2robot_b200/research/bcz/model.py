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
