"""Mock model and input generator for tests (utils/mocks.py:37-236): a linearly separable dataset and a 3-layer
feed-forward network with the reference's variable names (`MockT2RModel.dense.0/kernel`, `MockT2RModel.batch_norm.0/gamma`,
...), so the reference's own checkpoint fixture (test_data/mock_exported_savedmodel/variables) loads by name."""
import numpy as np
import torch

from tensor2robot_b200 import nn
from tensor2robot_b200.input_generators import abstract_input_generator
from tensor2robot_b200.models import abstract_model
from tensor2robot_b200.utils import dtypes
from tensor2robot_b200.utils import tensorspec_utils
from tensor2robot_b200.utils import tfdata

SEED = 1234
POSITIVE_SIZE = 64
TRAIN = 'train'


class MockInputGenerator(abstract_input_generator.AbstractInputGenerator):
  """Negative and positive samples with binary labels (mocks.py:44-96)."""

  def __init__(self, multi_dataset=False, **kwargs):
    self._multi_dataset = multi_dataset
    super(MockInputGenerator, self).__init__(**kwargs)

  def create_numpy_data(self):
    """A deterministic, linearly separable dataset: (features [128, 3], labels [128, 1])."""
    np.random.seed(SEED)
    positive = np.random.uniform(low=0.2, high=1.0, size=(POSITIVE_SIZE, 3))
    negative = np.random.uniform(low=-1.0, high=-0.2, size=(POSITIVE_SIZE, 3))
    features = np.concatenate([positive, negative], axis=0)
    labels = np.concatenate([np.ones((POSITIVE_SIZE, 1)), np.zeros((POSITIVE_SIZE, 1))], axis=0)
    return features, labels

  def _create_dataset(self, mode, params=None):
    batch_size = tfdata.get_batch_size(params, self._batch_size)
    features, labels = self.create_numpy_data()
    features, labels = features.astype(np.float32), labels.astype(np.float32)
    rng = np.random.RandomState(SEED)
    n = features.shape[0]
    while True:
      order = rng.permutation(n) if mode == TRAIN else np.arange(n)
      for start in range(0, n - batch_size + 1, batch_size):        # drop_remainder
        idx = order[start:start + batch_size]
        f = {'x1': features[idx], 'x2': features[idx]} if self._multi_dataset else {'x': features[idx]}
        yield f, {'y': labels[idx]}
      if mode != TRAIN:
        return


class MockT2RModel(abstract_model.AbstractT2RModel):
  """3 x (dense + elu + batch norm) -> dense(1) 'logit' (mocks.py:99-190)."""

  def __init__(self, multi_dataset=False, **kwargs):
    self._multi_dataset = multi_dataset
    super(MockT2RModel, self).__init__(**kwargs)

  def get_feature_specification(self, mode):
    del mode
    spec = tensorspec_utils.TensorSpecStruct()
    if self._multi_dataset:
      spec.x1 = tensorspec_utils.ExtendedTensorSpec(shape=(3,), dtype=dtypes.float32, name='measured_position',
                                                    dataset_key='dataset1')
      spec.x2 = tensorspec_utils.ExtendedTensorSpec(shape=(3,), dtype=dtypes.float32, name='measured_position',
                                                    dataset_key='dataset2')
    else:
      spec.x = tensorspec_utils.ExtendedTensorSpec(shape=(3,), dtype=dtypes.float32, name='measured_position')
    return spec

  def get_label_specification(self, mode):
    del mode
    spec = tensorspec_utils.TensorSpecStruct()
    spec.y = tensorspec_utils.ExtendedTensorSpec(shape=(1,), dtype=dtypes.float32, name='valid_position',
                                                 dataset_key='dataset1' if self._multi_dataset else '')
    return spec

  def inference_network_fn(self, features, labels, mode, config=None, params=None):
    del labels, mode, config, params
    net = (features.x1 + features.x2) if self._multi_dataset else features.x
    net = net.float()
    for pos, units in enumerate([32, 16, 8]):
      net = nn.elu(nn.dense_f32(net, units, scope='MockT2RModel.dense.{}'.format(pos), regularize=False,
                                names=('kernel', 'bias')))
      net = nn.batch_normalization_f32(net, 'MockT2RModel.batch_norm.{}'.format(pos))
    net = nn.dense_f32(net, 1, scope='MockT2RModel.dense.4', regularize=False, names=('kernel', 'bias'))
    return {'logit': net}

  def model_train_fn(self, features, labels, inference_outputs, mode, config=None, params=None):
    """tf.keras.losses.categorical_hinge: max(0, max((1 - y) * p) - sum(y * p) + 1), averaged over the batch."""
    del features, mode, config, params
    y_true, y_pred = labels.y.float(), inference_outputs['logit']
    pos = (y_true * y_pred).sum(-1)
    neg = ((1.0 - y_true) * y_pred).max(-1).values
    return torch.clamp(neg - pos + 1.0, min=0.0).mean()
