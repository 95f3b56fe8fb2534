"""The Tensor2Robot spec system on the B200 engine: ExtendedTensorSpec, TensorSpecStruct and the
spec algebra (flatten / pack / validate / copy / placeholders / feature dicts / assets).

This is the model-facing contract the reference defines in utils/tensorspec_utils.py; behaviour
(key order, optional handling, error types, proto wire format) follows that file and the
assertions of utils/tensorspec_utils_test.py, re-implemented over numpy/torch instead of
TensorFlow:

  ExtendedTensorSpec            utils/tensorspec_utils.py:40-278
  TensorSpecStruct              :302-682  (a flat ordered mapping with hierarchical *views*)
  flatten_spec_structure        :1303-1345 (namedtuple fields in order, dict keys sorted, like tf.nest)
  pack_flat_sequence_to_spec_structure  :1348-1427
  validate_and_pack / _flatten  :1210-1277
  assert_equal / assert_required / assert_valid_spec_structure  :1099-1207, :1463-1529
  tensorspec_to_feature_dict    :1571-1628
  pad_or_clip_tensor_to_spec_shape :1631-1682
  t2r_assets (de)serialisation  :1685-1733 and proto/t2r.proto

Spec/index work is exact: nothing here touches floating point values.
"""
import collections
import collections.abc
import contextlib
import logging
import pickle
import pprint

import numpy as np
import torch

from google.protobuf import text_format

from tensor2robot_b200.proto import t2r_pb2
from tensor2robot_b200.utils import dtypes

EXTRA_ASSETS_DIRECTORY = 'assets.extra'
T2R_ASSETS_FILENAME = 't2r_assets.pbtxt'


def _as_shape(shape):
  if shape is None:
    raise TypeError('shape must not be None')
  if isinstance(shape, (int, np.integer)):
    return (int(shape),)
  return tuple(None if d is None else int(d) for d in shape)


class TensorSpec(object):
  """Describes a tensor: shape, dtype, name (the role tf.TensorSpec plays in the reference)."""

  __slots__ = ('_shape', '_dtype', '_name')

  def __init__(self, shape, dtype, name=None):
    self._shape = _as_shape(shape)
    self._dtype = dtypes.as_dtype(dtype)
    self._name = name

  @property
  def shape(self):
    return self._shape

  @property
  def dtype(self):
    return self._dtype

  @property
  def name(self):
    return self._name

  def __eq__(self, other):
    return (isinstance(other, TensorSpec) and self._shape == other._shape and self._dtype == other._dtype)

  def __ne__(self, other):
    return not self == other

  def __hash__(self):
    return hash((self._shape, self._dtype))

  def __repr__(self):
    return 'TensorSpec(shape={}, dtype={}, name={})'.format(self.shape, repr(self.dtype), repr(self.name))


def _is_tensor(x):
  return isinstance(x, (torch.Tensor, np.ndarray))


class ExtendedTensorSpec(TensorSpec):
  """TensorSpec + is_optional / is_sequence / is_extracted / data_format / dataset_key /
  varlen_default_value (utils/tensorspec_utils.py:40-278)."""

  __slots__ = ('_is_optional', '_is_sequence', '_is_extracted', '_data_format', '_dataset_key',
               '_varlen_default_value')

  def __init__(self, shape, dtype, name=None, is_optional=None, is_sequence=False, is_extracted=False,
               data_format=None, dataset_key=None, varlen_default_value=None):
    super(ExtendedTensorSpec, self).__init__(shape, dtype, name)
    self._is_optional = False if is_optional is None else is_optional
    self._is_sequence = is_sequence
    self._is_extracted = is_extracted
    self._data_format = data_format
    self._dataset_key = '' if dataset_key is None else dataset_key
    self._varlen_default_value = varlen_default_value
    if varlen_default_value is not None:
      if data_format is None and len(self.shape) != 1:
        raise ValueError(('VarLenFeatures are only supported for shapes of rank 1 ({}) when '
                          'not using an image spec.').format(shape))
      if data_format is not None and len(self.shape) != 4:
        raise ValueError(('VarLenFeatures are only supported for shapes of rank 4 ({}) when '
                          'using an image spec.').format(shape))

  @classmethod
  def from_spec(cls, spec, shape=None, dtype=None, name=None, is_optional=None, is_sequence=None,
                is_extracted=None, data_format=None, dataset_key=None, batch_size=None,
                varlen_default_value=None):
    if not isinstance(spec, TensorSpec):
      raise ValueError('from_spec requires TensorSpec or ExtendedTensorSpec.')
    if is_optional is None:
      is_optional = getattr(spec, 'is_optional', False)
    if is_sequence is None:
      is_sequence = getattr(spec, 'is_sequence', False)
    if is_extracted is None:
      is_extracted = getattr(spec, 'is_extracted', False)
    if data_format is None:
      data_format = getattr(spec, 'data_format', None)
    if dataset_key is None:
      dataset_key = getattr(spec, 'dataset_key', '')
    shape = spec.shape if shape is None else _as_shape(shape)
    if batch_size:
      if not isinstance(batch_size, int):
        raise ValueError('batch_size must be an integer.')
      shape = ((None,) if batch_size < 0 else (batch_size,)) + tuple(shape)
    if varlen_default_value is None:
      varlen_default_value = getattr(spec, 'varlen_default_value', None)
    return cls(shape, dtype or spec.dtype, name or spec.name, is_optional, is_sequence, is_extracted,
               data_format, dataset_key, varlen_default_value)

  @classmethod
  def from_tensor(cls, tensor, name=None):
    if not _is_tensor(tensor):
      raise ValueError('`tensor` should be a torch.Tensor or np.ndarray')
    return cls(tuple(tensor.shape), dtypes.as_dtype(tensor.dtype), name, is_extracted=True)

  @classmethod
  def from_proto(cls, proto):
    kwargs = {'shape': tuple(proto.shape), 'dtype': dtypes.as_dtype(int(proto.dtype))}
    for field in ('name', 'is_optional', 'is_extracted', 'data_format', 'dataset_key', 'varlen_default_value'):
      if proto.HasField(field):
        kwargs[field] = getattr(proto, field)
    return cls(**kwargs)

  def to_proto(self):
    proto = t2r_pb2.ExtendedTensorSpec()
    proto.shape.extend([-1 if d is None else d for d in self.shape])
    proto.dtype = self.dtype.as_datatype_enum
    if self.name is not None:
      proto.name = self.name
    if self.is_optional is not None:
      proto.is_optional = self.is_optional
    if self.is_extracted is not None:
      proto.is_extracted = self.is_extracted
    if self.data_format is not None:
      proto.data_format = self.data_format
    if self.varlen_default_value is not None:
      proto.varlen_default_value = self.varlen_default_value
    return proto

  @classmethod
  def from_serialized_proto(cls, serialized):
    proto = t2r_pb2.ExtendedTensorSpec()
    proto.ParseFromString(serialized)
    return cls.from_proto(proto)

  @classmethod
  def to_spec(cls, instance):
    if isinstance(instance, TensorSpec):
      return cls.from_spec(instance)
    if _is_tensor(instance):
      return cls.from_tensor(instance)
    raise ValueError('We cannot convert {} with type {} to ExtendedTensorSpec'.format(instance, type(instance)))

  @property
  def is_optional(self):
    return self._is_optional

  @property
  def is_sequence(self):
    return self._is_sequence

  @property
  def is_extracted(self):
    return self._is_extracted

  @property
  def data_format(self):
    return self._data_format

  @property
  def dataset_key(self):
    return self._dataset_key

  @property
  def varlen_default_value(self):
    return self._varlen_default_value

  __hash__ = TensorSpec.__hash__

  def __repr__(self):
    return ('ExtendedTensorSpec(shape={}, dtype={}, name={}, is_optional={}, is_sequence={}, '
            'is_extracted={}, data_format={}, dataset_key={}, varlen_default_value={})').format(
                self.shape, repr(self.dtype), repr(self.name), repr(self.is_optional), repr(self.is_sequence),
                repr(self.is_extracted), repr(self.data_format), repr(self.dataset_key),
                repr(self.varlen_default_value))

  def __reduce__(self):
    return ExtendedTensorSpec, (self._shape, self._dtype, self._name, self._is_optional, self._is_sequence,
                                self._is_extracted, self._data_format, self._dataset_key,
                                self._varlen_default_value)


def _namedtuple_to_dict(item):
  if isinstance(item, tuple) and hasattr(item, '_asdict'):
    return collections.OrderedDict(item._asdict())
  return item


class TensorSpecStruct(collections.abc.MutableMapping):
  """A flat ordered {path: spec | tensor} mapping with hierarchical attribute *views*.

  `struct.train` is a view on every key below 'train/'; views share storage with the struct they
  were taken from, so adding/deleting through a view is visible everywhere
  (utils/tensorspec_utils.py:302-682 and the behaviour pinned by tensorspec_utils_test.py:154-283).
  """

  def __init__(self, *args, **kwargs):
    object.__setattr__(self, '_prefix', kwargs.pop('__internal_path_prefix', ''))
    object.__setattr__(self, '_root', kwargs.pop('__internal_dict_view', None) or collections.OrderedDict())
    if len(args) > 1:
      raise TypeError('expected at most 1 positional argument, got %d' % len(args))
    if args:
      init = args[0]
      items = init.items() if isinstance(init, collections.abc.Mapping) else init
      for key, value in items:
        self[key] = value
    for key, value in kwargs.items():
      if not key.startswith('_'):
        self[key] = value

  # -- path helpers ---------------------------------------------------------------------------
  def _full(self, key):
    return self._prefix + '/' + key if self._prefix else key

  def _own_keys(self):
    if not self._prefix:
      return list(self._root.keys())
    start = self._prefix + '/'
    return [k[len(start):] for k in self._root.keys() if k.startswith(start)]

  # -- mapping protocol -----------------------------------------------------------------------
  def __getitem__(self, key):
    full = self._full(key)
    if full in self._root:
      return self._root[full]
    start = full + '/'
    if any(k.startswith(start) for k in self._root):
      return TensorSpecStruct(__internal_path_prefix=full, __internal_dict_view=self._root)
    raise AttributeError('No attribute with the name {} exists for {}'.format(key, self))

  def __setitem__(self, key, value):
    value = _namedtuple_to_dict(value)
    if isinstance(value, TensorSpecStruct) or isinstance(value, dict):
      if not value:
        raise ValueError('We cannot assign an empty TensorSpecStruct or dict. Please, first fill it.')
      for sub_key, sub_value in list(value.items()):
        self[key + '/' + sub_key] = sub_value
      return
    if value is not None and not isinstance(value, TensorSpec) and not _is_tensor(value):
      raise ValueError('Only TensorSpecs, tensors, np.ndarrays, None or non-empty dict-like structures of '
                       'them can be assigned; got {} for {}.'.format(type(value), key))
    self._root[self._full(key)] = value

  def __delitem__(self, key):
    full = self._full(key)
    if full in self._root:
      del self._root[full]
      return
    start = full + '/'
    sub = [k for k in self._root if k.startswith(start)]
    if not sub:
      raise KeyError(key)
    for k in sub:
      del self._root[k]

  def __iter__(self):
    return iter(self._own_keys())

  def __len__(self):
    return len(self._own_keys())

  def __contains__(self, key):
    return isinstance(key, str) and self._full(key) in self._root

  def keys(self):
    return self._own_keys()

  def values(self):
    return [self._root[self._full(k)] for k in self._own_keys()]

  def items(self):
    return [(k, self._root[self._full(k)]) for k in self._own_keys()]

  # -- attribute access -----------------------------------------------------------------------
  def __getattr__(self, name):
    if name.startswith('_'):
      raise AttributeError('The attribute {} does not exist.'.format(name))
    return self[name]

  def __setattr__(self, name, value):
    if name.startswith('_'):
      object.__setattr__(self, name, value)
    else:
      self[name] = value

  def __delattr__(self, name):
    try:
      del self[name]
    except KeyError:
      raise AttributeError(name)

  def __repr__(self):
    return 'TensorSpecStruct(\n' + pprint.pformat(self.to_dict()) + ')'

  def __reduce__(self):
    return TensorSpecStruct, (list(self.items()),)

  def __copy__(self):
    return TensorSpecStruct(list(self.items()))

  def __deepcopy__(self, memo):
    import copy
    return TensorSpecStruct([(k, copy.deepcopy(v, memo)) for k, v in self.items()])

  def to_dict(self):
    """A new shallow dict of this view."""
    return dict(self.items())

  # -- protos ---------------------------------------------------------------------------------
  @classmethod
  def from_proto(cls, proto):
    return cls(sorted((k, ExtendedTensorSpec.from_proto(v)) for k, v in proto.key_value.items()))

  @classmethod
  def from_serialized_proto(cls, serialized):
    proto = t2r_pb2.TensorSpecStruct()
    proto.ParseFromString(serialized)
    return cls.from_proto(proto)

  def to_proto(self):
    proto = t2r_pb2.TensorSpecStruct()
    for key, value in self.items():
      if not hasattr(value, 'to_proto'):
        raise ValueError('Only data structures which support to_proto, e.g. ExtendedTensorSpec are allowed '
                         'within a TensorSpecStruct when converting to a proto. The type for key {} is '
                         'however {} with the value {}.'.format(key, type(value), value))
      proto.key_value[key].CopyFrom(value.to_proto())
    return proto


# ---------------------------------------------------------------------------------------------
# structure traversal (what tf.nest does in the reference)
# ---------------------------------------------------------------------------------------------
def _is_leaf(value):
  return value is None or isinstance(value, TensorSpec) or _is_tensor(value)


def _children(structure):
  """[(key, child)] in tf.nest order: namedtuple fields in order, mapping keys sorted, sequences by index."""
  if isinstance(structure, TensorSpecStruct):
    return list(structure.items())       # already flat and ordered
  if isinstance(structure, tuple) and hasattr(structure, '_fields'):
    return [(f, getattr(structure, f)) for f in structure._fields]
  if isinstance(structure, collections.abc.Mapping):
    return [(k, structure[k]) for k in sorted(structure.keys())]
  if isinstance(structure, (list, tuple)):
    return [(str(i), v) for i, v in enumerate(structure)]
  raise ValueError('We only support spec_structures of (hierarchical) dicts or namedtuples, not {}.'.format(
      type(structure)))


def _flatten_with_paths(structure, prefix=''):
  out = []
  for key, child in _children(structure):
    path = prefix + '/' + key if prefix else key
    if _is_leaf(child):
      out.append((path, child))
    else:
      out.extend(_flatten_with_paths(child, path))
  return out


def _map_structure(fn, structure):
  if _is_leaf(structure):
    return fn(structure)
  if isinstance(structure, TensorSpecStruct):
    return TensorSpecStruct([(k, fn(v)) for k, v in structure.items()])
  if isinstance(structure, tuple) and hasattr(structure, '_fields'):
    return type(structure)(*[_map_structure(fn, getattr(structure, f)) for f in structure._fields])
  if isinstance(structure, collections.OrderedDict):
    return collections.OrderedDict((k, _map_structure(fn, v)) for k, v in structure.items())
  if isinstance(structure, collections.abc.Mapping):
    return type(structure)((k, _map_structure(fn, v)) for k, v in structure.items())
  if isinstance(structure, (list, tuple)):
    return type(structure)(_map_structure(fn, v) for v in structure)
  raise ValueError('unsupported structure type {}'.format(type(structure)))


def convert_to_tensorspecstruct(inputs):
  return TensorSpecStruct(inputs)


def is_flat_spec_or_tensors_structure(spec_or_tensors):
  """True for {key: spec | tensor} mappings without nesting (utils/tensorspec_utils.py:1430-1460)."""
  if isinstance(spec_or_tensors, (collections.abc.Mapping, TensorSpecStruct)):
    for value in spec_or_tensors.values():
      if isinstance(value, TensorSpec) or _is_tensor(value):
        continue
      return False
    return True
  return False


def assert_valid_spec_structure(spec_structure, used_tensorspec_names=None):
  """Hierarchies of dicts / namedtuples / lists of specs or tensors only; specs sharing a `name` must
  agree in shape and dtype (utils/tensorspec_utils.py:1463-1529)."""
  if used_tensorspec_names is None:
    used_tensorspec_names = {}
  spec_structure = _namedtuple_to_dict(spec_structure)
  if isinstance(spec_structure, (collections.abc.Mapping, TensorSpecStruct)):
    values = list(spec_structure.values())
  elif isinstance(spec_structure, (list, tuple)):
    values = list(spec_structure)
  else:
    raise ValueError('We only support spec_structures of (hierarchical) dicts or namedtuples, not {}.'.format(
        type(spec_structure)))
  for value in values:
    if isinstance(value, TensorSpec):
      if value.name is not None:
        if value.name in used_tensorspec_names:
          try:
            assert_equal_spec_or_tensor(used_tensorspec_names[value.name], value)
          except ValueError:
            raise ValueError(
                'All TensorSpecs with a name defined have to be unique or non unique specs have to define '
                'the same shape and dtype within our spec_structure. Yet, the name {} is defined twice and '
                'describes different specs {} vs {}.'.format(value.name, value,
                                                             used_tensorspec_names[value.name]))
        used_tensorspec_names[value.name] = value
      continue
    if _is_tensor(value) or value is None:
      continue
    assert_valid_spec_structure(value, used_tensorspec_names)


def flatten_spec_structure(spec_structure, filter_none=True):
  """Flat TensorSpecStruct keyed by '/'-joined paths (utils/tensorspec_utils.py:1303-1345)."""
  assert_valid_spec_structure(spec_structure)
  if is_flat_spec_or_tensors_structure(spec_structure) and isinstance(spec_structure, TensorSpecStruct):
    items = list(spec_structure.items())
  elif is_flat_spec_or_tensors_structure(spec_structure):
    items = [(k, spec_structure[k]) for k in spec_structure.keys()]
  else:
    items = _flatten_with_paths(spec_structure)
  if filter_none:
    items = [(k, v) for k, v in items if v is not None]
  out = TensorSpecStruct()
  for k, v in items:
    out[k] = v
  return out


def pack_flat_sequence_to_spec_structure(spec_structure, flat_sequence_with_joined_string_paths):
  """Packs a flat {path: tensor} mapping into the shape of `spec_structure`; absent optional specs
  become None, absent required specs raise (utils/tensorspec_utils.py:1348-1427)."""
  assert_valid_spec_structure(spec_structure)
  if not is_flat_spec_or_tensors_structure(flat_sequence_with_joined_string_paths):
    raise ValueError('The provided flat_sequence_with_joined_string_paths is not a flat sequence {}.'.format(
        flat_sequence_with_joined_string_paths))
  flat = dict(flat_sequence_with_joined_string_paths.items())

  def lookup(path, spec):
    if path in flat:
      return flat[path]
    if spec is None or getattr(spec, 'is_optional', False):
      if spec is not None:
        logging.info('The optional TensorSpec %s is not present at %s.', spec, path)
      return None
    raise ValueError('The required {} spec {} is not available.'.format(path, spec))

  if isinstance(spec_structure, TensorSpecStruct) or is_flat_spec_or_tensors_structure(spec_structure):
    flat_spec = flatten_spec_structure(spec_structure, filter_none=False)
    out = TensorSpecStruct()
    for key in sorted(flat_spec.keys()):        # reference sorts TensorSpecStruct keys (:1384-1391)
      value = lookup(key, flat_spec[key])
      if value is not None:
        out[key] = value
    return out

  def pack(structure, prefix):
    if _is_leaf(structure):
      return lookup(prefix, structure)
    if isinstance(structure, tuple) and hasattr(structure, '_fields'):
      return type(structure)(*[pack(getattr(structure, f), prefix + '/' + f if prefix else f)
                               for f in structure._fields])
    if isinstance(structure, collections.abc.Mapping):
      ctor = collections.OrderedDict if isinstance(structure, collections.OrderedDict) else dict
      return ctor((k, pack(structure[k], prefix + '/' + k if prefix else k)) for k in structure.keys())
    if isinstance(structure, (list, tuple)):
      return type(structure)(pack(v, prefix + '/' + str(i) if prefix else str(i)) for i, v in enumerate(structure))
    raise ValueError('unsupported structure type {}'.format(type(structure)))

  return pack(spec_structure, '')


def maybe_ignore_batch(spec_or_tensors, ignore_batch=False):
  """Strips the leading (batch) dimension of every leaf (utils/tensorspec_utils.py:1072-1096)."""
  if not ignore_batch:
    return spec_or_tensors

  def map_fn(spec):
    if spec is None:
      return None
    if _is_tensor(spec):
      return ExtendedTensorSpec(tuple(spec.shape[1:]), dtypes.as_dtype(spec.dtype), None, is_extracted=True)
    return ExtendedTensorSpec.from_spec(spec, shape=spec.shape[1:])
  return _map_structure(map_fn, spec_or_tensors)


_DEVICE_BF16_AS_F32 = [False]


@contextlib.contextmanager
def device_bf16_as_float32():
  """Inside this context a CUDA bfloat16 *tensor* satisfies a float32 spec.

  The engine stores activations in bf16 on the device while model specs stay float32, the same
  arrangement the reference makes for TPUs (preprocessors/tpu_preprocessor_wrapper.py rewrites
  float32 specs to bfloat16, models/tpu_model_wrapper.py casts back)."""
  _DEVICE_BF16_AS_F32.append(True)
  try:
    yield
  finally:
    _DEVICE_BF16_AS_F32.pop()


def assert_equal_spec_or_tensor(expected_spec_or_tensor, actual_spec_or_tensor):
  """dtype, rank and every non-None dimension must agree (utils/tensorspec_utils.py:1099-1139)."""
  expected_spec = ExtendedTensorSpec.to_spec(expected_spec_or_tensor)
  actual_spec = ExtendedTensorSpec.to_spec(actual_spec_or_tensor)
  if expected_spec.is_sequence and actual_spec.is_extracted:
    actual_spec = maybe_ignore_batch(actual_spec, ignore_batch=True)
  device_bf16 = (_DEVICE_BF16_AS_F32[-1] and expected_spec.dtype == dtypes.float32 and
                 actual_spec.dtype == dtypes.bfloat16 and actual_spec.is_extracted)
  if expected_spec.dtype != actual_spec.dtype and not device_bf16:
    raise ValueError('TensorSpec.dtype {} does not match TensorSpec.dtype {} in specs\n expected: {}\n actual: {}'
                     .format(expected_spec.dtype, actual_spec.dtype, expected_spec, actual_spec))
  if len(expected_spec.shape) != len(actual_spec.shape):
    raise ValueError('TensorSpec.shape {} does not match TensorSpec.shape {} in specs\n expected: {}\n actual: {}'
                     .format(expected_spec.shape, actual_spec.shape, expected_spec, actual_spec))
  for expected_dim, actual_dim in zip(expected_spec.shape, actual_spec.shape):
    if expected_dim is None:
      continue
    if expected_dim != actual_dim:
      raise ValueError('TensorSpec.shape {} does not match TensorSpec.shape {}.'.format(
          expected_spec.shape, actual_spec.shape))


def assert_equal(expected_tensors_or_spec, actual_tensors_or_spec, ignore_batch=False):
  """Same structure, shapes and dtypes (utils/tensorspec_utils.py:1142-1166)."""
  actual_tensors_or_spec = maybe_ignore_batch(actual_tensors_or_spec, ignore_batch)
  expected = flatten_spec_structure(expected_tensors_or_spec)
  actual = flatten_spec_structure(actual_tensors_or_spec)
  if sorted(expected.keys()) != sorted(actual.keys()):
    raise ValueError('The two structures don\'t have the same keys: {} vs {}'.format(
        sorted(expected.keys()), sorted(actual.keys())))
  for key in expected.keys():
    assert_equal_spec_or_tensor(expected[key], actual[key])


def assert_required(expected_spec, actual_tensors_or_spec, ignore_batch=False):
  """Every required spec of `expected_spec` is present in `actual` and matches
  (utils/tensorspec_utils.py:1169-1207)."""
  flat_actual = flatten_spec_structure(actual_tensors_or_spec)
  packed = pack_flat_sequence_to_spec_structure(expected_spec, flat_actual)
  flat_actual = flatten_spec_structure(packed)
  flat_expected = flatten_spec_structure(expected_spec)
  flat_expected = {k: v for k, v in flat_expected.items() if k in flat_actual}
  assert_equal(flat_expected, flat_actual, ignore_batch)


def _log_mismatch(expected_spec, actual_tensors_or_spec):
  logging.error('The actual_spec_or_tensor does not fulfill the expected_spec:')
  for key, value in sorted(flatten_spec_structure(expected_spec).items()):
    logging.error('expected_spec: %s: %s', key, value)
  for key, value in sorted(flatten_spec_structure(actual_tensors_or_spec).items()):
    logging.error('actual_spec:   %s: %s', key, ExtendedTensorSpec.to_spec(value) if _is_tensor(value) else value)


def validate_and_flatten(expected_spec, actual_tensors_or_spec, ignore_batch=False):
  """utils/tensorspec_utils.py:1210-1241."""
  assert_valid_spec_structure(expected_spec)
  assert_valid_spec_structure(actual_tensors_or_spec)
  try:
    assert_required(expected_spec, actual_tensors_or_spec, ignore_batch)
  except ValueError:
    _log_mismatch(expected_spec, actual_tensors_or_spec)
    raise
  return flatten_spec_structure(actual_tensors_or_spec)


def validate_and_pack(expected_spec, actual_tensors_or_spec, ignore_batch=False):
  """utils/tensorspec_utils.py:1244-1277."""
  assert_valid_spec_structure(expected_spec)
  assert_valid_spec_structure(actual_tensors_or_spec)
  if not is_flat_spec_or_tensors_structure(actual_tensors_or_spec):
    actual_tensors_or_spec = flatten_spec_structure(actual_tensors_or_spec)
  try:
    assert_required(expected_spec, actual_tensors_or_spec, ignore_batch)
  except ValueError:
    _log_mismatch(expected_spec, actual_tensors_or_spec)
    raise
  return pack_flat_sequence_to_spec_structure(expected_spec, actual_tensors_or_spec)


def add_sequence_length_specs(spec_structure):
  """Adds a key + '_length' int64 scalar spec for every sequence spec (:1280-1288)."""
  flat = flatten_spec_structure(spec_structure)
  for key, value in flat.items():
    if value.is_sequence:
      flat[key + '_length'] = ExtendedTensorSpec(shape=(), dtype=dtypes.int64, name=value.name + '_length')
  return flat


def filter_spec_structure_by_dataset(spec_structure, dataset_key, filter_none=True):
  """Subset of the flattened structure whose dataset_key matches (:1291-1300)."""
  flat = flatten_spec_structure(spec_structure, filter_none)
  return TensorSpecStruct([(k, v) for k, v in flat.items() if (v.dataset_key == dataset_key or not dataset_key)])


def filter_required_flat_tensor_spec(flat_tensor_spec):
  """Drops optional specs from a flat structure (:1532-1555)."""
  if not is_flat_spec_or_tensors_structure(flat_tensor_spec):
    raise ValueError('Only flat tensor_spec structures are allowed.')
  out = TensorSpecStruct()
  for key, value in flat_tensor_spec.items():
    if hasattr(value, 'is_optional') and value.is_optional:
      continue
    out[key] = value
  return out


def replace_dtype(tensor_spec_struct, from_dtype, to_dtype):
  for key, value in tensor_spec_struct.items():
    if value.dtype == from_dtype:
      tensor_spec_struct[key] = ExtendedTensorSpec.from_spec(spec=value, dtype=to_dtype)
  return tensor_spec_struct


def cast_float32_to_bfloat16(tensor_spec_struct, output_spec):
  for key, value in output_spec.items():
    if value is not None and value.dtype == dtypes.bfloat16:
      if dtypes.as_dtype(tensor_spec_struct[key].dtype) != dtypes.float32:
        raise ValueError('Attempting to convert non float32 type {} to bfloat16 for the element {} with the '
                         'name {}.'.format(tensor_spec_struct[key].dtype, tensor_spec_struct[key], key))
      tensor_spec_struct[key] = torch.as_tensor(tensor_spec_struct[key]).to(torch.bfloat16)
  return tensor_spec_struct


def cast_bfloat16_to_float32(tensor_spec_struct):
  for key, value in tensor_spec_struct.items():
    if value is not None and dtypes.as_dtype(value.dtype) == dtypes.bfloat16:
      tensor_spec_struct[key] = value.to(torch.float32)
  return tensor_spec_struct


def copy_tensorspec(spec_structure, prefix='', batch_size=None):
  """Copy with names re-prefixed and an optional batch dimension (:755-780)."""
  assert_valid_spec_structure(spec_structure)
  if prefix:
    prefix += '/'

  def map_spec(spec):
    if spec is None:
      return None
    name = spec.name or ''
    return ExtendedTensorSpec.from_spec(spec, name=prefix + name, batch_size=batch_size)
  return _map_structure(map_spec, spec_structure)


def _batched_shape(t, batch_size, sequence_length=None):
  shape = tuple(t.shape)
  if sequence_length is not None and getattr(t, 'is_sequence', False):
    shape = (sequence_length,) + shape
  if batch_size is None:
    shape = (None,) + shape
  elif batch_size > 0:
    shape = (batch_size,) + shape
  return shape


def make_placeholders(spec_structure, batch_size=None):
  """The reference returns tf.placeholders (:783-814); without a graph the equivalent is the batched
  spec itself: same structure, shape (batch,)+shape (None = variable)."""
  assert_valid_spec_structure(spec_structure)

  def make_placeholder(t):
    t = ExtendedTensorSpec.from_spec(t)
    shape = tuple(t.shape)
    if t.is_sequence:
      shape = (None,) + shape
    if batch_size is None:
      shape = (None,) + shape
    elif batch_size > 0:
      shape = (batch_size,) + shape
    return ExtendedTensorSpec.from_spec(t, shape=shape)
  return _map_structure(make_placeholder, spec_structure)


def make_constant_numpy(spec_structure, constant_value, batch_size=2, sequence_length=3):
  """:847-883."""
  assert_valid_spec_structure(spec_structure)

  def make_fixed(t):
    shape = _batched_shape(t, batch_size, sequence_length)
    return np.full(shape, constant_value).astype(t.dtype.as_numpy_dtype)
  return _map_structure(make_fixed, spec_structure)


def make_random_numpy(spec_structure, batch_size=2, sequence_length=3):
  """uniform[0,255) for uint8/int32/int64 specs, [0,1) otherwise, cast to the spec dtype (:886-920)."""
  assert_valid_spec_structure(spec_structure)

  def make_random(t):
    maxval = 255 if t.dtype in (dtypes.uint8, dtypes.int32, dtypes.int64) else 1.0
    shape = _batched_shape(t, batch_size, sequence_length)
    r = np.random.uniform(size=shape, high=maxval)
    return r.astype(t.dtype.as_numpy_dtype)
  return _map_structure(make_random, spec_structure)


def make_random_tensors(spec_structure, batch_size=2):
  """:817-844, as torch CPU tensors."""
  return _map_structure(lambda a: torch.from_numpy(a), make_random_numpy(spec_structure, batch_size, None))


def map_feed_dict_unsafe(feature_placeholders_spec, np_inputs_spec):
  """Deprecated (:1012-1045): {placeholder key: numpy} without checking dtypes / shapes / unused inputs; numpy inputs
  that the spec does not know are dropped with a warning."""
  logging.warning('map_feed_dict_unsafe is deprecated. Please update to map_feed_dict.')
  flat_spec = flatten_spec_structure(feature_placeholders_spec)
  flat_np_inputs = flatten_spec_structure(np_inputs_spec)
  for key in flat_np_inputs.keys():
    if key not in flat_spec:
      logging.warning('np_inputs has an input: %s, not found in the tensorspec.', key)
  return {key: flat_np_inputs[key] for key in flat_spec.keys()}      # a missing input fails on lookup, as in the reference


def map_feed_dict(spec_placeholders, spec_numpy, feed_dict=None, ignore_batch=False):
  """{placeholder key: numpy} after validating the arrays against the specs (:923-965)."""
  if not is_flat_spec_or_tensors_structure(spec_placeholders):
    spec_placeholders = flatten_spec_structure(spec_placeholders)
  if not is_flat_spec_or_tensors_structure(spec_numpy):
    spec_numpy = flatten_spec_structure(spec_numpy)
  if feed_dict is None:
    feed_dict = {}
  assert_required(maybe_ignore_batch(spec_placeholders, ignore_batch), maybe_ignore_batch(spec_numpy, ignore_batch))
  for key, value in spec_numpy.items():
    if key in feed_dict:
      raise ValueError('We would overwrite existing placeholder mapping {}.'.format(key))
    feed_dict[key] = value
  return feed_dict


def map_predict_fn_dict(spec_structure, spec_numpy, feed_dict=None, ignore_batch=False):
  """:968-1009."""
  if not is_flat_spec_or_tensors_structure(spec_numpy):
    spec_numpy = flatten_spec_structure(spec_numpy)
  if feed_dict is None:
    feed_dict = {}
  assert_required(spec_structure, maybe_ignore_batch(spec_numpy, ignore_batch))
  for key, value in spec_numpy.items():
    if key not in spec_structure:
      continue
    if key in feed_dict:
      raise ValueError('We would overwrite existing placeholder mapping {}.'.format(key))
    feed_dict[key] = value
  return feed_dict


def tensorspec_from_tensors(tensors):
  """Spec structure inferred from a structure of tensors (:1043-1069)."""
  assert_valid_spec_structure(tensors)
  return _map_structure(lambda t: None if t is None else ExtendedTensorSpec.to_spec(t), tensors)


# ---------------------------------------------------------------------------------------------
# tf.Example feature descriptions
# ---------------------------------------------------------------------------------------------
FixedLenFeature = collections.namedtuple('FixedLenFeature', ['shape', 'dtype', 'default_value'])
FixedLenSequenceFeature = collections.namedtuple('FixedLenSequenceFeature', ['shape', 'dtype', 'allow_missing'])
VarLenFeature = collections.namedtuple('VarLenFeature', ['dtype'])


def is_encoded_image_spec(tensor_spec):
  """data_format 'jpeg' / 'png' (case-insensitive) marks an encoded image (:1558-1568)."""
  if hasattr(tensor_spec, 'data_format'):
    return tensor_spec.data_format is not None and tensor_spec.data_format.upper() in ['JPEG', 'PNG']
  return False


def _get_feature(tensor_spec, decode_images=True):
  """:1571-1593."""
  varlen_default_value = getattr(tensor_spec, 'varlen_default_value', None)
  if getattr(tensor_spec, 'is_sequence', False):
    cls = lambda shape, dtype: FixedLenSequenceFeature(shape, dtype, True)
  elif varlen_default_value is not None:
    cls = lambda shape, dtype: VarLenFeature(dtype)
  else:
    cls = lambda shape, dtype: FixedLenFeature(shape, dtype, None)
  if decode_images and is_encoded_image_spec(tensor_spec):
    if varlen_default_value is not None:
      return cls((), dtypes.string)
    if len(tensor_spec.shape) > 3:
      return cls((tensor_spec.shape[0],), dtypes.string)
    return cls((), dtypes.string)
  return cls(tuple(tensor_spec.shape), tensor_spec.dtype)


def tensorspec_to_feature_dict(tensor_spec_struct, decode_images=True):
  """({tf.Example key: feature description}, {key: spec}); specs without a name are not parsed
  (:1596-1628)."""
  assert_valid_spec_structure(tensor_spec_struct)
  features = {}
  tensor_spec_dict = {}
  for key, tensor_spec in flatten_spec_structure(tensor_spec_struct).items():
    if tensor_spec.name is None:
      logging.info('TensorSpec name attribute for %s is not set; will not parse this Tensor from TFExamples.', key)
      continue
    features[tensor_spec.name] = _get_feature(tensor_spec, decode_images)
    tensor_spec_dict[tensor_spec.name] = tensor_spec
  return features, tensor_spec_dict


def pad_or_clip_tensor_to_spec_shape(tensor, tensor_spec):
  """[B, N, ...] -> [B, T, ...] with T = spec.shape[0]: right-pad with varlen_default_value or clip
  (:1631-1682).  `tensor` may be a dense array or a list of per-row arrays (ragged)."""
  target = tensor_spec.shape[0]
  np_dtype = tensor_spec.dtype.as_numpy_dtype
  default = np.asarray(tensor_spec.varlen_default_value).astype(np_dtype)
  rows = [np.asarray(r) for r in tensor]
  rest = tuple(tensor_spec.shape[1:])
  out = np.full((len(rows), target) + rest, default, dtype=np_dtype)
  for i, r in enumerate(rows):
    r = r.astype(np_dtype).reshape((-1,) + rest)
    n = min(r.shape[0], target)
    out[i, :n] = r[:n]
  return out


# ---------------------------------------------------------------------------------------------
# assets
# ---------------------------------------------------------------------------------------------
def write_t2r_assets_to_file(t2r_assets, filename):
  with open(filename, 'w') as f:
    f.write(text_format.MessageToString(t2r_assets))


def load_t2r_assets_to_file(filename):
  with open(filename, 'r') as f:
    t2r_assets = t2r_pb2.T2RAssets()
    text_format.Parse(f.read(), t2r_assets)
    return t2r_assets


def write_input_spec_to_file(in_feature_spec, in_label_spec, filename):
  with open(filename, 'wb') as f:
    pickle.dump({'in_feature_spec': in_feature_spec, 'in_label_spec': in_label_spec}, f)


def load_input_spec_from_file(filename):
  import os
  if not os.path.exists(filename):
    raise ValueError('The file {} does not exist.'.format(filename))
  with open(filename, 'rb') as f:
    spec_data = pickle.load(f)
  return spec_data['in_feature_spec'], spec_data['in_label_spec']


def write_global_step_to_file(global_step, filename):
  with open(filename, 'wb') as f:
    pickle.dump({'global_step': global_step}, f)


def load_global_step_from_file(filename):
  import os
  if not os.path.exists(filename):
    raise ValueError('The file {} does not exist.'.format(filename))
  with open(filename, 'rb') as f:
    return pickle.load(f)['global_step']
