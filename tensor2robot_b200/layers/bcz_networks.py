"""BC-Z policy heads (layers/bcz_networks.py:107-145)."""
import torch

from tensor2robot_b200 import nn


def _relu32(x):
  return nn.relu(x)


def MultiHeadMLP(net, action_sizes, num_waypoints, fc_layers, is_training,  # pylint: disable=invalid-name
                 stop_gradient_future_waypoints=True):
  """One (fc_layers..., Linear(action_size * num_waypoints)) MLP per action component on the fp32
  policy features `net` [B, F]; returns a list of [B, num_waypoints, action_size] tensors.  With
  num_waypoints > 1 the first waypoint comes from the `action_trajectory` heads and the remaining
  ones from `auxiliary_trajectory` heads whose input gradient is stopped in training
  (bcz_networks.py:107-145).  Variable names follow slim's layers.stack / fully_connected numbering."""
  if net.dim() != 2:
    raise NotImplementedError('MultiHeadMLP over [B, T, F] sequences is not built')

  def mlp(x, waypoints):
    outputs = []
    counter = {'stack': 0, 'fc': 0}

    def name(kind):
      n = counter[kind]
      counter[kind] += 1
      base = 'Stack' if kind == 'stack' else 'fully_connected'
      return base if n == 0 else '%s_%d' % (base, n)

    for action_size in action_sizes:
      head = x
      with nn.variable_scope(name('stack')):
        for i, units in enumerate(fc_layers):
          head = _relu32(nn.dense_f32(head, units, scope='fully_connected_%d' % (i + 1)))
      head = nn.dense_f32(head, action_size * waypoints, scope=name('fc'))
      outputs.append(head.reshape(-1, waypoints, action_size))
    return outputs

  if num_waypoints > 1 and stop_gradient_future_waypoints:
    with nn.variable_scope('action_trajectory'):
      components_1 = mlp(net, 1)
    with nn.variable_scope('auxiliary_trajectory'):
      components_2 = mlp(net.detach() if is_training else net, num_waypoints - 1)
    return [torch.cat([c1, c2], dim=-2) for c1, c2 in zip(components_1, components_2)]
  return mlp(net, num_waypoints)
