"""Episode data -> transition Examples for the pose toy env (research/pose_env/episode_to_transitions.py:31-51): the
writer of the fixture format the pose_env models train on."""
from PIL import Image

from tensor2robot_b200.utils import example_proto
from tensor2robot_b200.utils import image


def episode_to_transitions_pose_toy(episode_data):
  """episode_data: [(obs_t uint8 [64,64,3], action, reward, obs_tp1, done, debug)] -> [Example].  A supervised
  regression problem: obs_tp1 and done are dropped."""
  transitions = []
  for obs_t, action, reward, _obs_tp1, _done, debug in episode_data:
    features = {}
    features['state/image'] = example_proto.bytes_feature([image.jpeg_string(Image.fromarray(obs_t))])
    features['pose'] = example_proto.float_feature(action.flatten().tolist())
    features['reward'] = example_proto.float_feature([reward])
    features['target_pose'] = example_proto.float_feature(debug['target_pose'].tolist())
    transitions.append(example_proto.Example(features=features))
  return transitions
