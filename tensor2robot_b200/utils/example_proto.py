"""tf.train.Example wire-format *encoder* for the replay-writing side (tensorflow/core/example/{example,feature}.proto;
the reference builds these messages in research/pose_env/episode_to_transitions.py:22-48 and serialises them in
utils/writer.py:49-61).  TensorFlow / protobuf-generated classes are not available, so the few messages are encoded
directly: Example{features = 1}, Features{map<string, Feature> feature = 1}, Feature{bytes_list = 1 | float_list = 2 |
int64_list = 3}, the lists hold `repeated value = 1` (floats and int64s packed).  Map entries are written in insertion
order (protobuf defines no canonical map order; the runtime that wrote the reference's fixture varies it per record) -
re-inserting the parsed features in each record's wire order reproduces that fixture byte for byte (tests/test_writer.py)."""

import numpy as np


def _varint(value):
  value &= (1 << 64) - 1            # negative int64 -> 10-byte two's complement varint
  out = bytearray()
  while True:
    b = value & 0x7F
    value >>= 7
    if value:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _ld(field, payload):
  """Length-delimited field."""
  return _varint((field << 3) | 2) + _varint(len(payload)) + payload


class Feature(object):
  """tf.train.Feature with exactly one of the three lists."""

  def __init__(self, kind, values):
    self.kind = kind          # 'bytes' | 'float' | 'int64'
    self.values = values

  def SerializeToString(self):  # pylint: disable=invalid-name
    if self.kind == 'bytes':
      return _ld(1, b''.join(_ld(1, bytes(v)) for v in self.values))
    if self.kind == 'float':
      packed = np.asarray(self.values, dtype='<f4').tobytes()
      return _ld(2, _ld(1, packed) if packed else b'')
    packed = b''.join(_varint(int(v)) for v in self.values)
    return _ld(3, _ld(1, packed) if packed else b'')


def bytes_feature(values):
  if isinstance(values, (bytes, bytearray)):
    values = [values]
  return Feature('bytes', [bytes(v) for v in values])


def float_feature(values):
  return Feature('float', [float(v) for v in np.asarray(values, dtype=np.float64).ravel()])


def int64_feature(values):
  return Feature('int64', [int(v) for v in np.asarray(values).ravel()])


class Example(object):
  """tf.train.Example(features=tf.train.Features(feature={...})): `features` maps key -> Feature."""

  def __init__(self, features=None):
    self.features = dict(features or {})

  def SerializeToString(self):  # pylint: disable=invalid-name
    entries = b''.join(_ld(1, _ld(1, key.encode('utf-8')) + _ld(2, feature.SerializeToString()))
                       for key, feature in self.features.items())
    return _ld(1, entries)


class SequenceExample(object):
  """tf.train.SequenceExample{context = 1, feature_lists = 2}: feature_lists maps key -> list of Feature."""

  def __init__(self, context=None, feature_lists=None):
    self.context = dict(context or {})
    self.feature_lists = dict(feature_lists or {})

  def SerializeToString(self):  # pylint: disable=invalid-name
    context = b''.join(_ld(1, _ld(1, k.encode('utf-8')) + _ld(2, f.SerializeToString())) for k, f in self.context.items())
    lists = b''
    for key, features in self.feature_lists.items():
      feature_list = b''.join(_ld(1, f.SerializeToString()) for f in features)
      lists += _ld(1, _ld(1, key.encode('utf-8')) + _ld(2, feature_list))
    return (_ld(1, context) if self.context else b'') + (_ld(2, lists) if self.feature_lists else b'')
