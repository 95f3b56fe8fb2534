"""Spec-system parity: the assertions of the reference's utils/tensorspec_utils_test.py
(:75-82 flat order, :101-108 required/optional, :154-283 struct views, :134-144 proto round trip,
:389-431 pack ordering, :696-722 pad_or_clip) re-expressed against tensor2robot_b200."""
import collections
import copy
import os
import pickle

import numpy as np
import pytest

from tensor2robot_b200.utils import dtypes
from tensor2robot_b200.utils import tensorspec_utils as utils

TSPEC = utils.ExtendedTensorSpec
T1 = TSPEC((224, 224, 3), dtypes.float32, 'images', data_format='jpeg')
T2 = TSPEC((6), dtypes.float32, 'actions')
T3 = TSPEC((), dtypes.float32, 'reward')
O4 = TSPEC((224, 224, 3), dtypes.float32, 'debug_images', is_optional=True)
O6 = TSPEC((6), dtypes.float32, 'debug_actions', is_optional=True)
S7 = TSPEC((6), dtypes.float32, 'sequence_actions', is_sequence=True)
D1 = TSPEC((224, 224, 3), dtypes.float32, 'debug_images', dataset_key='d1')
D2 = TSPEC((224, 224, 3), dtypes.float32, 'debug_images', dataset_key='d2')

MockBar = collections.namedtuple('Bar', ['images', 'actions'])
MockNested = collections.namedtuple('Nested', ['train', 'test'])
MockNestedOptional = collections.namedtuple('NestedOptional', ['train', 'test', 'optional'])
MockNestedSubset = collections.namedtuple('NestedSubset', ['train'])
nested = MockNested(train=MockBar(T1, T2), test=MockBar(T1, T2))
nested_optional = MockNestedOptional(train=MockBar(T1, T2), test=MockBar(T1, T2), optional=MockBar(O4, O6))
nested_subset = MockNestedSubset(train=MockBar(T1, T2))

REFERENCE_FLAT = collections.OrderedDict([
    ('train/images', T1), ('train/actions', T2), ('test/images', T1), ('test/actions', T2),
    ('optional/images', O4), ('optional/actions', O6)])


def test_flatten_order_and_values():
  flat = utils.flatten_spec_structure(nested_optional)
  assert list(flat.keys()) == list(REFERENCE_FLAT.keys())
  assert flat.to_dict() == dict(REFERENCE_FLAT)
  assert utils.flatten_spec_structure(nested_subset).to_dict() == {'train/images': T1, 'train/actions': T2}
  # plain dict keys are visited sorted, like tf.nest
  assert list(utils.flatten_spec_structure({'b': T2, 'a': {'z': T1, 'y': T3}}).keys()) == ['a/y', 'a/z', 'b']


def test_assert_equal_and_required():
  utils.assert_equal(nested, copy.deepcopy(nested))
  with pytest.raises(ValueError):
    utils.assert_equal(nested, nested_subset)
  utils.assert_required(nested_subset, nested)          # a subset is asked for
  utils.assert_required(nested_optional, nested)        # extras are optional
  with pytest.raises(ValueError):
    utils.assert_required(nested, nested_subset)        # more required specs than available


def test_init_with_attributes():
  train = utils.TensorSpecStruct(images=T1, actions=T2)
  flat = utils.flatten_spec_structure(nested_optional)
  utils.assert_equal(train, flat.train)
  alternative = {'o6': O6, 'o4': O4}
  hierarchy = utils.TensorSpecStruct(nested_optional_spec=nested_optional, alternative=alternative)
  utils.assert_equal(hierarchy.nested_optional_spec, flat)
  assert hierarchy.alternative.to_dict() == alternative
  assert sorted(hierarchy.keys()) == sorted(
      ['nested_optional_spec/' + k for k in REFERENCE_FLAT] + ['alternative/o6', 'alternative/o4'])


def test_proto_round_trips(tmp_path):
  t1 = utils.ExtendedTensorSpec.from_serialized_proto(T1.to_proto().SerializeToString())
  utils.assert_equal_spec_or_tensor(T1, t1)
  assert t1.name == 'images' and t1.data_format == 'jpeg'
  struct = utils.TensorSpecStruct(REFERENCE_FLAT)
  back = utils.TensorSpecStruct.from_serialized_proto(struct.to_proto().SerializeToString())
  assert struct.to_dict() == back.to_dict()
  assert back['optional/images'].is_optional
  # is_sequence is not serialised (reference :177-208)
  assert not utils.ExtendedTensorSpec.from_proto(S7.to_proto()).is_sequence
  from tensor2robot_b200.proto import t2r_pb2
  assets = t2r_pb2.T2RAssets()
  assets.feature_spec.CopyFrom(struct.to_proto())
  assets.global_step = 7
  path = os.path.join(str(tmp_path), utils.T2R_ASSETS_FILENAME)
  utils.write_t2r_assets_to_file(assets, path)
  loaded = utils.load_t2r_assets_to_file(path)
  assert loaded.global_step == 7
  assert utils.TensorSpecStruct.from_proto(loaded.feature_spec).to_dict() == struct.to_dict()


def test_reference_asset_file_is_readable():
  """The reference's own t2r_assets.pbtxt (copied verbatim as a fixture) parses: dtype enum 1 = float32."""
  here = os.path.dirname(os.path.abspath(__file__))
  assets = utils.load_t2r_assets_to_file(os.path.join(here, 'golden', 'mock_t2r_assets.pbtxt'))
  feature_spec = utils.TensorSpecStruct.from_proto(assets.feature_spec)
  assert feature_spec.x.shape == (3,) and feature_spec.x.dtype == dtypes.float32
  assert feature_spec.x.name == 'measured_position'
  assert utils.TensorSpecStruct.from_proto(assets.label_spec).y.name == 'valid_position'


def test_struct_views():
  s = utils.TensorSpecStruct(REFERENCE_FLAT)
  assert s.to_dict() == dict(REFERENCE_FLAT)
  assert list(s.train.keys()) == ['images', 'actions']
  s.train.addition = O6                                   # propagates to the parent
  assert list(s.train.keys()) == ['images', 'actions', 'addition']
  assert list(s.keys()) == list(REFERENCE_FLAT.keys()) + ['train/addition']
  assert s['train/addition'] is s.train.addition
  with pytest.raises(AttributeError):
    _ = s['optional_typo']
  with pytest.raises(AttributeError):
    _ = s.optional_typo
  assert s['optional'].to_dict() == s.optional.to_dict()
  del s['optional/images']                                # deletion propagates down ...
  with pytest.raises(AttributeError):
    _ = s.optional.images
  test = s.test
  del test['actions']                                     # ... and up
  assert 'test/actions' not in s
  assert utils.TensorSpecStruct(s.to_dict()).to_dict() == s.to_dict()


def test_struct_assignment_rules():
  s = utils.TensorSpecStruct()
  with pytest.raises(ValueError):
    s.should_raise = utils.TensorSpecStruct()
  with pytest.raises(ValueError):
    s.should_raise = {}
  with pytest.raises(ValueError):
    s.should_raise = 'a string'
  sub = utils.TensorSpecStruct()
  sub.data = np.ones(1)
  s.sub_data = sub
  s.sub_data.additional = np.zeros(1)
  assert list(s.keys()) == ['sub_data/data', 'sub_data/additional']
  # prefix handling: 'val_mode' is not below 'val'
  t = utils.TensorSpecStruct()
  t.val_mode = np.ones(1)
  val = utils.TensorSpecStruct()
  val.mode = np.zeros(1)
  val.data = np.ones(1)
  t.val = val
  assert list(t.val.keys()) == ['mode', 'data'] and list(t['val'].keys()) == ['mode', 'data']


def test_struct_composition():
  s = utils.TensorSpecStruct(REFERENCE_FLAT)
  new_field = utils.TensorSpecStruct(REFERENCE_FLAT)
  s.new_field = new_field
  assert list(s.new_field.keys()) == list(new_field.keys())
  assert list(s.keys()) == list(REFERENCE_FLAT.keys()) + ['new_field/' + k for k in new_field.keys()]
  test_spec = MockNested(train={'a': np.ones(1), 'b': 2 * np.ones(1)}, test=MockBar(T1, T2))
  ref = utils.flatten_spec_structure(test_spec)
  new = utils.flatten_spec_structure(MockNested(train=ref.train, test=MockBar(T1, T2)))
  for key in ref:
    assert key in new
    assert new[key] is ref[key]


def test_filter_required():
  flat = utils.flatten_spec_structure(nested_optional)
  required = utils.filter_required_flat_tensor_spec(flat)
  assert required.to_dict() == {'train/images': T1, 'train/actions': T2, 'test/images': T1, 'test/actions': T2}
  with pytest.raises(ValueError):
    utils.filter_required_flat_tensor_spec(nested_optional)


def test_tensorspec_to_feature_dict():
  features, spec_dict = utils.tensorspec_to_feature_dict(nested_subset, decode_images=True)
  assert spec_dict == {'images': T1, 'actions': T2}
  assert features == {'images': utils.FixedLenFeature((), dtypes.string, None),
                      'actions': utils.FixedLenFeature(T2.shape, T2.dtype, None)}
  features, _ = utils.tensorspec_to_feature_dict(nested_subset, decode_images=False)
  assert features['images'] == utils.FixedLenFeature(T1.shape, T1.dtype, None)
  # specs without a name are not parsed (reference :1620-1625)
  features, _ = utils.tensorspec_to_feature_dict({'a': TSPEC((1,), dtypes.float32)})
  assert features == {}


def test_assert_equal_spec_or_tensor():
  utils.assert_equal_spec_or_tensor(T1, T1)
  utils.assert_equal_spec_or_tensor(T1, TSPEC((224, 224, 3), dtypes.float32, name='random'))
  utils.assert_equal_spec_or_tensor(T1, np.zeros((224, 224, 3), np.float32))
  for bad in (T2, TSPEC((224, 223, 3), dtypes.float32), TSPEC((224, 224, 3), dtypes.uint8),
              np.zeros((224, 224, 3), np.uint8)):
    with pytest.raises(ValueError):
      utils.assert_equal_spec_or_tensor(T1, bad)


def test_is_flat():
  assert not utils.is_flat_spec_or_tensors_structure(nested_subset)
  assert utils.is_flat_spec_or_tensors_structure(utils.flatten_spec_structure(nested_subset))
  assert not utils.is_flat_spec_or_tensors_structure([T1, T2])
  assert utils.is_flat_spec_or_tensors_structure({'t1': T1, 't2': T2})


def test_pack_flat_sequence():
  subset_ph = utils.make_placeholders(nested_subset)
  packed = utils.pack_flat_sequence_to_spec_structure(nested_subset, utils.flatten_spec_structure(subset_ph))
  utils.assert_equal(subset_ph, packed)
  utils.assert_equal(nested_subset, packed, ignore_batch=True)
  ph = utils.make_placeholders(nested)
  flat_ph = utils.flatten_spec_structure(ph)
  packed = utils.pack_flat_sequence_to_spec_structure(nested_subset, flat_ph)
  utils.assert_equal(nested_subset, packed, ignore_batch=True)
  packed_optional = utils.pack_flat_sequence_to_spec_structure(nested_optional, flat_ph)
  assert packed_optional.optional.images is None        # optional and absent
  utils.assert_required(packed_optional, ph)
  utils.assert_required(nested, packed_optional, ignore_batch=True)
  with pytest.raises(ValueError):
    utils.pack_flat_sequence_to_spec_structure(nested, utils.flatten_spec_structure(subset_ph))


def test_pack_orders_struct_keys():
  spec = utils.TensorSpecStruct()
  spec.b = TSPEC((1,), dtypes.float32, 'b')
  spec.a = TSPEC((1,), dtypes.float32, 'a')
  spec.c = TSPEC((1,), dtypes.float32, 'c')
  packed = utils.pack_flat_sequence_to_spec_structure(spec, utils.make_placeholders(spec))
  assert list(packed.keys()) == ['a', 'b', 'c']
  assert [v.name for v in packed.values()] == ['a', 'b', 'c']


def test_validate_flatten_and_pack():
  features = utils.make_random_numpy(nested, batch_size=2)
  flat = utils.validate_and_flatten(nested_optional, features, ignore_batch=True)
  packed = utils.validate_and_pack(nested_subset, flat, ignore_batch=True)
  assert isinstance(packed, MockNestedSubset)          # packed back into the expected structure
  assert packed.train.images.shape == (2, 224, 224, 3)
  packed_struct = utils.validate_and_pack(utils.flatten_spec_structure(nested_subset), flat, ignore_batch=True)
  assert list(packed_struct.keys()) == ['train/actions', 'train/images']
  bad = utils.make_random_numpy(nested_subset, batch_size=2)
  with pytest.raises(ValueError):
    utils.validate_and_pack(nested, bad, ignore_batch=True)


def test_sequence_and_dataset_helpers():
  modified = utils.add_sequence_length_specs(utils.TensorSpecStruct(image1=D1, actions=S7))
  assert modified.actions_length == TSPEC((), dtypes.int64, 'sequence_actions_length')
  assert modified.actions_length.name == 'sequence_actions_length'
  spec = utils.TensorSpecStruct(image1=D1, image2=D2)
  for key, name, d in zip(['d1', 'd2'], ['image1', 'image2'], [D1, D2]):
    assert utils.filter_spec_structure_by_dataset(spec, key).to_dict() == {name: d}


def test_valid_and_invalid_structures():
  for structure in ({'a': T1, 'b': T2}, MockBar(T1, T2), [T1, T2], (T1, T2)):
    utils.assert_valid_spec_structure(structure)
  with pytest.raises(ValueError):
    utils.assert_valid_spec_structure({'test': 10})
  with pytest.raises(ValueError):
    utils.assert_valid_spec_structure(MockBar(images=TSPEC((3, 2, 3), dtypes.float32, 'images'),
                                              actions=TSPEC((3, 2), dtypes.float32, 'images')))
  # the same name twice with the same shape/dtype is fine (reference :1503-1515)
  utils.assert_valid_spec_structure(MockBar(images=T1, actions=TSPEC((224, 224, 3), dtypes.float32, 'images')))


def test_copy_and_placeholders():
  copied = utils.copy_tensorspec(nested_subset, prefix='p', batch_size=4)
  assert copied.train.images.name == 'p/images' and copied.train.images.shape == (4, 224, 224, 3)
  assert utils.copy_tensorspec(nested_subset, batch_size=-1).train.actions.shape == (None, 6)
  ph = utils.make_placeholders(utils.TensorSpecStruct(seq=S7, x=T2), batch_size=None)
  assert ph.seq.shape == (None, None, 6) and ph.x.shape == (None, 6)
  const = utils.make_constant_numpy(utils.TensorSpecStruct(seq=S7, x=T2), 3.0, batch_size=2, sequence_length=5)
  assert const.seq.shape == (2, 5, 6) and float(const.x[0, 0]) == 3.0
  rnd = utils.make_random_numpy({'u': TSPEC((4,), dtypes.uint8, 'u'), 'f': TSPEC((4,), dtypes.float32, 'f')}, 3)
  assert rnd['u'].dtype == np.uint8 and rnd['f'].dtype == np.float32 and rnd['f'].max() < 1.0


def test_varlen_spec_rules_and_pad_or_clip():
  with pytest.raises(ValueError):
    TSPEC((2, 3), dtypes.float32, 'x', varlen_default_value=1.0)
  with pytest.raises(ValueError):
    TSPEC((2, 3, 4), dtypes.uint8, 'x', data_format='png', varlen_default_value=1.0)
  spec = TSPEC((3,), dtypes.float32, 'varlen', varlen_default_value=3.0)
  out = utils.pad_or_clip_tensor_to_spec_shape([[1], [1, 2]], spec)
  np.testing.assert_array_equal(out, [[1, 3, 3], [1, 2, 3]])          # reference :696-712
  out = utils.pad_or_clip_tensor_to_spec_shape([[1, 2, 3, 4]], spec)
  np.testing.assert_array_equal(out, [[1, 2, 3]])                      # reference :714-722


def test_pickle_and_eq():
  assert pickle.loads(pickle.dumps(T1)) == T1
  assert pickle.loads(pickle.dumps(T1)).data_format == 'jpeg'
  s = utils.TensorSpecStruct(REFERENCE_FLAT)
  assert pickle.loads(pickle.dumps(s)).to_dict() == s.to_dict()
  assert T1 == TSPEC((224, 224, 3), dtypes.float32, 'other')     # equality ignores the name (:261-263)
  assert dtypes.as_dtype(np.float32) == dtypes.float32 and dtypes.as_dtype(1) == dtypes.float32


def test_stem_weight_layouts_round_trip():
  """nn._stem_pack / _stem_unpack: [Cout,KH,KW,3] <-> the K layouts of t2r_stem_conv_* (include/t2r_b200.h):
  one filter row per 64-wide chunk, or two rows x 8 pixels for stride-2 stems with KW <= 8."""
  import numpy as np
  from tensor2robot_b200 import nn
  rng = np.random.RandomState(0)
  for kh, kw, stride, k_expected in ((7, 7, 2, 256), (6, 6, 2, 192), (7, 7, 1, 448), (3, 9, 2, 192)):
    w = rng.standard_normal((64, kh, kw, 3)).astype(np.float32)
    packed = nn._stem_pack(w, stride)
    assert packed.shape == (64, k_expected)
    np.testing.assert_array_equal(nn._stem_unpack(packed, kh, kw, 3, stride), w)
    assert np.count_nonzero(packed) == np.count_nonzero(w)     # padding slots are zero
  # rows == 2 layout: slot = px * 8 + row * 4 + channel inside chunk kh // 2
  w = np.zeros((1, 7, 7, 3), np.float32)
  w[0, 3, 5, 2] = 1.0
  packed = nn._stem_pack(w, 2)
  assert packed[0, (3 // 2) * 64 + 5 * 8 + (3 % 2) * 4 + 2] == 1.0
