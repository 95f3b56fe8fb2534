"""Data types of the spec system.

The reference's specs carry `tf.DType`s; TensorFlow is not part of this engine, so a small DType
registry stands in for it.  Each DType knows its numpy and torch equivalents and the TensorFlow
`DataType` enum value used on the wire by proto/t2r.proto (`ExtendedTensorSpec.dtype`), which keeps
`t2r_assets.pbtxt` files interchangeable with the reference.
"""
import numpy as np
import torch


class DType(object):
  """A named data type: `float32`, `uint8`, `string`, `bfloat16`, ..."""

  __slots__ = ('name', 'np_dtype', 'torch_dtype', 'as_datatype_enum')

  def __init__(self, name, np_dtype, torch_dtype, enum):
    self.name = name
    self.np_dtype = np_dtype
    self.torch_dtype = torch_dtype
    self.as_datatype_enum = enum

  @property
  def as_numpy_dtype(self):
    if self.np_dtype is None:
      raise TypeError('%s has no numpy equivalent' % self.name)
    return self.np_dtype

  @property
  def is_floating(self):
    return self.name in ('float16', 'bfloat16', 'float32', 'float64')

  @property
  def is_integer(self):
    return self.name in ('int8', 'int16', 'int32', 'int64', 'uint8', 'uint16', 'uint32', 'uint64')

  def __eq__(self, other):
    try:
      return self.name == as_dtype(other).name
    except TypeError:
      return False

  def __ne__(self, other):
    return not self == other

  def __hash__(self):
    return hash(self.name)

  def __repr__(self):
    return 'dtypes.%s' % self.name

  def __reduce__(self):
    return as_dtype, (self.name,)


float16 = DType('float16', np.float16, torch.float16, 19)
bfloat16 = DType('bfloat16', None, torch.bfloat16, 14)
float32 = DType('float32', np.float32, torch.float32, 1)
float64 = DType('float64', np.float64, torch.float64, 2)
int8 = DType('int8', np.int8, torch.int8, 6)
int16 = DType('int16', np.int16, torch.int16, 5)
int32 = DType('int32', np.int32, torch.int32, 3)
int64 = DType('int64', np.int64, torch.int64, 9)
uint8 = DType('uint8', np.uint8, torch.uint8, 4)
uint16 = DType('uint16', np.uint16, getattr(torch, 'uint16', None), 17)
uint32 = DType('uint32', np.uint32, getattr(torch, 'uint32', None), 22)
uint64 = DType('uint64', np.uint64, getattr(torch, 'uint64', None), 23)
bool_ = DType('bool', np.bool_, torch.bool, 10)
string = DType('string', np.object_, None, 7)

_ALL = [float16, bfloat16, float32, float64, int8, int16, int32, int64, uint8, uint16, uint32, uint64, bool_,
        string]
_BY_NAME = {d.name: d for d in _ALL}
_BY_ENUM = {d.as_datatype_enum: d for d in _ALL}
_BY_TORCH = {d.torch_dtype: d for d in _ALL if d.torch_dtype is not None}


def as_dtype(value):
  """Converts a DType, name, TF enum int, numpy dtype/type or torch dtype to a DType."""
  if isinstance(value, DType):
    return value
  if isinstance(value, str):
    if value in _BY_NAME:
      return _BY_NAME[value]
    if value in ('str', 'bytes', 'object'):
      return string
  if isinstance(value, bool):
    raise TypeError('cannot convert %r to a DType' % (value,))
  if isinstance(value, int) and value in _BY_ENUM:
    return _BY_ENUM[value]
  if isinstance(value, torch.dtype):
    if value in _BY_TORCH:
      return _BY_TORCH[value]
    raise TypeError('unsupported torch dtype %s' % value)
  if value in (str, bytes, object):
    return string
  try:
    npd = np.dtype(value)
  except TypeError:
    raise TypeError('cannot convert %r to a DType' % (value,))
  if npd.kind in 'OSU':
    return string
  if npd.name in _BY_NAME:
    return _BY_NAME[npd.name]
  raise TypeError('unsupported dtype %r' % (value,))
