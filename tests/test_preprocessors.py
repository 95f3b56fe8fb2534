"""NoOpPreprocessor with the cases of the reference's preprocessors/noop_preprocessor_test.py:28-197: hierarchical
namedtuple specs, optional members, flattened and packed inputs, labels optional, broken requirements."""
import collections

import numpy as np
import pytest
import torch

from tensor2robot_b200.preprocessors import noop_preprocessor
from tensor2robot_b200.utils import dtypes
from tensor2robot_b200.utils import tensorspec_utils

TSPEC = tensorspec_utils.ExtendedTensorSpec
MockFeatures = collections.namedtuple('MockFeatures', ['images', 'actions', 'optional_hierarchy'])
MockHierachy = collections.namedtuple('MockHierachy', ['debug_images'])
MockFeaturesRequired = collections.namedtuple('MockFeaturesRequired', ['images', 'actions'])
MockFeaturesBroken = collections.namedtuple('MockFeaturesBroken', ['images', 'actions', 'broken'])
MockLabels = collections.namedtuple('MockLabels', ['score'])

mock_features_required = MockFeaturesRequired(images=TSPEC(shape=(224, 224, 3), dtype=dtypes.float32),
                                              actions=TSPEC(shape=(6,), dtype=dtypes.float32))
mock_features_broken = MockFeaturesBroken(images=TSPEC(shape=(224, 224, 3), dtype=dtypes.float32),
                                          actions=TSPEC(shape=(6,), dtype=dtypes.float32),
                                          broken=TSPEC(shape=(1,), dtype=dtypes.float32))
mock_features = MockFeatures(
    images=mock_features_required.images, actions=mock_features_required.actions,
    optional_hierarchy=MockHierachy(debug_images=TSPEC(shape=(224, 224, 3), dtype=dtypes.float32, is_optional=True)))
mock_labels = MockLabels(score=TSPEC(shape=(1,), dtype=dtypes.float32))
mock_features_fn = lambda mode: mock_features
mock_features_broken_fn = lambda mode: mock_features_broken
mock_labels_fn = lambda mode: mock_labels


def test_init_noop_preprocessor():
  noop_preprocessor.NoOpPreprocessor(mock_features_fn, mock_labels_fn)


@pytest.mark.parametrize('spec_or_tensors', [
    {'test': 1},                                                                      # non_tensorspec_tensor_values_dict
    MockFeaturesRequired(images=np.random.random_sample(10), actions='action'),       # ..._named_tuple
])
def test_init_noop_preprocessor_raises(spec_or_tensors):
  spec_or_tensors_fn = lambda _: spec_or_tensors
  with pytest.raises(ValueError):
    noop_preprocessor.NoOpPreprocessor(spec_or_tensors_fn, mock_labels_fn)
  with pytest.raises(ValueError):
    noop_preprocessor.NoOpPreprocessor(mock_features_fn, spec_or_tensors_fn)


def _preprocess(preprocessor, feature_spec, label_spec, flatten):
  """Feeds random tensors shaped by the given specs (batch 1) and checks that they pass through unchanged."""
  np_features = tensorspec_utils.make_random_numpy(feature_spec, batch_size=1)
  np_labels = tensorspec_utils.make_random_numpy(label_spec, batch_size=1) if label_spec is not None else None
  to_torch = lambda struct: type(struct)(*[to_torch(v) if isinstance(v, tuple) else torch.from_numpy(v) for v in struct])
  features = to_torch(np_features)
  labels = to_torch(np_labels) if np_labels is not None else None
  if flatten:
    features = tensorspec_utils.flatten_spec_structure(features)
    labels = tensorspec_utils.flatten_spec_structure(labels) if labels is not None else None
  out_features, out_labels = preprocessor.preprocess(features=features, labels=labels, mode='train')
  for key, value in tensorspec_utils.flatten_spec_structure(np_features).items():
    np.testing.assert_allclose(value, out_features[key].numpy())
  if np_labels is not None:
    for key, value in tensorspec_utils.flatten_spec_structure(np_labels).items():
      np.testing.assert_allclose(value, out_labels[key].numpy())
  else:
    assert out_labels is None


def test_noop_preprocessor_preprocess_fn():
  preprocessor = noop_preprocessor.NoOpPreprocessor(mock_features_fn, mock_labels_fn)
  _preprocess(preprocessor, mock_features, mock_labels, flatten=True)       # everything, flattened and packed
  _preprocess(preprocessor, mock_features, mock_labels, flatten=False)
  _preprocess(preprocessor, mock_features_required, mock_labels, flatten=False)   # the optional member is absent
  _preprocess(preprocessor, mock_features_required, mock_labels, flatten=True)
  _preprocess(preprocessor, mock_features_required, None, flatten=True)           # labels are not required
  broken = noop_preprocessor.NoOpPreprocessor(mock_features_broken_fn, mock_labels_fn)
  with pytest.raises(ValueError):
    _preprocess(broken, mock_features_required, mock_labels, flatten=False)
  with pytest.raises(ValueError):
    _preprocess(broken, mock_features_required, mock_labels, flatten=True)


def test_abstract_preprocessor_is_abstract():
  """preprocessors/abstract_preprocessor_test.py."""
  from tensor2robot_b200.preprocessors import abstract_preprocessor
  with pytest.raises(TypeError):
    abstract_preprocessor.AbstractPreprocessor()        # pylint: disable=abstract-class-instantiated
