"""Grasp2Vec T2R model (research/grasp2vec/grasp2vec_model.py:40-240): unsupervised object embeddings from
(pregrasp, postgrasp, goal) image triples; scene images share one tower pass (2B images), the goal image has
its own tower; n-pairs loss between (pre - post) and goal."""
import numpy as np
import torch

from tensor2robot_b200.models import abstract_model
from tensor2robot_b200.preprocessors import image_transformations
from tensor2robot_b200.preprocessors import spec_transformation_preprocessor
from tensor2robot_b200.research.grasp2vec import losses
from tensor2robot_b200.research.grasp2vec import networks
from tensor2robot_b200.utils import dtypes
from tensor2robot_b200.utils import tensorspec_utils

TRAIN, EVAL, PREDICT = 'train', 'eval', 'infer'
TensorSpec = tensorspec_utils.ExtendedTensorSpec
_RNG = np.random.RandomState(0)


def seed(value):
  global _RNG
  _RNG = np.random.RandomState(value)


def maybe_crop_images(images, params, mode):
  """The same crop for every tensor of the list: random offsets in TRAIN, the interval midpoints otherwise
  (grasp2vec_model.py:44-72).  images: uint8 [B,H,W,3] CUDA tensors; returns views."""
  (min_offset_height, max_offset_height, target_height, min_offset_width, max_offset_width, target_width) = params
  if mode == TRAIN:
    offset_height = int(_RNG.randint(min_offset_height, max_offset_height))   # maxval exclusive like tf
    offset_width = int(_RNG.randint(min_offset_width, max_offset_width))
  else:
    offset_height = (min_offset_height + max_offset_height) // 2
    offset_width = (min_offset_width + max_offset_width) // 2
  images = [img[:, offset_height:offset_height + target_height, offset_width:offset_width + target_width]
            for img in images]
  return images, offset_height, offset_width


def _random_flips(image):
  """tf.image.random_flip_left_right / _up_down on a 4-D batch: one coin per IMAGE and per axis
  (TF 1.15 image_ops_impl._random_flip, rank-4 branch)."""
  b = image.shape[0]
  for dim in (2, 1):
    coins = torch.from_numpy(_RNG.uniform(size=b) > 0.5).to(image.device)
    image = torch.where(coins.view(b, 1, 1, 1), torch.flip(image, dims=[dim]), image)
  return image


class Grasp2VecPreprocessor(spec_transformation_preprocessor.SpecTransformationPreprocessor):
  """Crop, convert to [0, 1], random flips in TRAIN (grasp2vec_model.py:76-133)."""

  def __init__(self, scene_crop=(0, 40, 472, 0, 168, 472), goal_crop=(0, 40, 472, 0, 168, 472), **kwargs):
    self._scene_crop = scene_crop
    self._goal_crop = goal_crop
    super(Grasp2VecPreprocessor, self).__init__(**kwargs)

  def _transform_in_feature_specification(self, flat_spec_structure):
    for name in ['pregrasp_image', 'postgrasp_image', 'goal_image']:
      self.update_spec(flat_spec_structure, name, shape=(512, 640, 3), dtype=dtypes.uint8, data_format='jpeg')
    return flat_spec_structure

  def _preprocess_fn(self, features, labels, mode):
    scene_images, _, _ = maybe_crop_images([features['pregrasp_image'], features['postgrasp_image']],
                                           self._scene_crop, mode)
    features['pregrasp_image'] = scene_images[0]
    features['postgrasp_image'] = scene_images[1]
    features['goal_image'] = maybe_crop_images([features['goal_image']], self._goal_crop, mode)[0][0]
    for name in ['pregrasp_image', 'postgrasp_image', 'goal_image']:
      image = image_transformations.convert_and_distort(features[name], None)   # uint8 crop view -> bf16 [0, 1]
      if mode == TRAIN:
        image = _random_flips(image)
      features[name] = image
    return features, labels


class Grasp2VecModel(abstract_model.AbstractT2RModel):
  """Basic Grasp2Vec model."""

  def __init__(self, scene_size, goal_size, embedding_loss_fn=losses.NPairsLoss, global_negatives=False, **kwargs):
    """global_negatives=True (data-parallel runs, SURVEY 8e opt-in): the n-pairs loss draws its negatives from every
    replica's batch through an all-gather of the [B, 1024] embeddings; the default keeps the reference's per-replica
    loss."""
    self._scene_size = tuple(scene_size)
    self._goal_size = tuple(goal_size)
    self._embedding_loss_fn = embedding_loss_fn
    self._global_negatives = global_negatives
    super(Grasp2VecModel, self).__init__(**kwargs)

  def get_feature_specification(self, mode):
    tspec = tensorspec_utils.TensorSpecStruct()
    tspec.pregrasp_image = TensorSpec(shape=self._scene_size + (3,), dtype=dtypes.float32, name='image',
                                      data_format='jpeg')
    tspec.postgrasp_image = TensorSpec(shape=self._scene_size + (3,), dtype=dtypes.float32, name='postgrasp_image',
                                       data_format='jpeg')
    tspec.goal_image = TensorSpec(shape=self._goal_size + (3,), dtype=dtypes.float32, name='present_image',
                                  data_format='jpeg')
    return tspec

  def get_label_specification(self, mode):
    return tensorspec_utils.TensorSpecStruct()   # Grasp2Vec is unsupervised

  @property
  def default_preprocessor_cls(self):
    return Grasp2VecPreprocessor

  def inference_network_fn(self, features, labels, mode, config=None, params=None):
    """Scene images run as ONE 2B batch through the `scene` tower (grasp2vec_model.py:180-203)."""
    scene_images = torch.cat([features.pregrasp_image, features.postgrasp_image], dim=0)
    v, s = networks.Embedding(scene_images, mode, params, scope='scene')
    b = features.pregrasp_image.shape[0]
    goal_v, goal_s = networks.Embedding(features.goal_image, mode, params, scope='goal')
    return {'pre_vector': v[:b], 'post_vector': v[b:], 'pre_spatial': s[:b], 'post_spatial': s[b:],
            'goal_vector': goal_v, 'goal_spatial': goal_s}

  def model_train_fn(self, features, labels, inference_outputs, mode, config=None, params=None):
    kwargs = {'global_negatives': True} if self._global_negatives else {}
    embed_loss = self._embedding_loss_fn(inference_outputs['pre_vector'], inference_outputs['goal_vector'],
                                         inference_outputs['post_vector'], **kwargs)
    return embed_loss, {'embed_loss': embed_loss}
