"""ORACLE (test infrastructure, never imported by the product): pure-Python restatement of the
on-disk formats of the record path.

  TFRecord framing + masked CRC-32C     SURVEY Appendix B; what tf.data.TFRecordDataset reads
                                        (reference call site utils/tfdata.py:174-210)
  tf.Example / Feature wire format      what tf.parse_example decodes (utils/tfdata.py:385)

The format definitions live in TensorFlow (core/lib/io/record_writer.cc, core/example/example.proto),
a dependency absent from /root/reference; PINNED against the reference's own fixture
test_data/pose_env_test_data.tfrecord (copied to tests/golden/) and, when importable, against
tensorboard's independent masked_crc32c.
"""
import struct

_POLY = 0x82F63B78
_TABLE = []
for _i in range(256):
  _c = _i
  for _ in range(8):
    _c = (_c >> 1) ^ _POLY if _c & 1 else _c >> 1
  _TABLE.append(_c)


def crc32c(data):
  crc = 0xFFFFFFFF
  for b in data:
    crc = (crc >> 8) ^ _TABLE[(crc ^ b) & 0xFF]
  return crc ^ 0xFFFFFFFF


def masked_crc32c(data):
  crc = crc32c(data)
  return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF


def read_tfrecords(path, verify=True):
  out = []
  with open(path, 'rb') as f:
    data = f.read()
  pos = 0
  while pos < len(data):
    (length,) = struct.unpack_from('<Q', data, pos)
    (len_crc,) = struct.unpack_from('<I', data, pos + 8)
    if verify and masked_crc32c(data[pos:pos + 8]) != len_crc:
      raise ValueError('length crc mismatch at %d' % pos)
    payload = data[pos + 12:pos + 12 + length]
    (data_crc,) = struct.unpack_from('<I', data, pos + 12 + length)
    if verify and masked_crc32c(payload) != data_crc:
      raise ValueError('data crc mismatch at %d' % pos)
    out.append(payload)
    pos += 12 + length + 4
  return out


def _varint(buf, pos):
  result = shift = 0
  while True:
    b = buf[pos]
    pos += 1
    result |= (b & 0x7F) << shift
    if not b & 0x80:
      return result, pos
    shift += 7


def _fields(buf):
  pos = 0
  while pos < len(buf):
    tag, pos = _varint(buf, pos)
    field, wire = tag >> 3, tag & 7
    if wire == 0:
      value, pos = _varint(buf, pos)
    elif wire == 1:
      value, pos = buf[pos:pos + 8], pos + 8
    elif wire == 2:
      n, pos = _varint(buf, pos)
      value, pos = buf[pos:pos + n], pos + n
    elif wire == 5:
      value, pos = buf[pos:pos + 4], pos + 4
    else:
      raise ValueError('unsupported wire type %d' % wire)
    yield field, wire, value


def parse_example(serialized):
  """{key: ('bytes'|'float'|'int64', [values])} of one serialized tf.Example."""
  out = {}
  for field, wire, features in _fields(serialized):
    if field != 1 or wire != 2:
      continue
    for f2, w2, entry in _fields(features):
      if f2 != 1 or w2 != 2:
        continue
      key, feature = None, b''
      for f3, _, value in _fields(entry):
        if f3 == 1:
          key = value.decode('utf-8')
        elif f3 == 2:
          feature = value
      kind, values = None, []
      for f4, _, lst in _fields(feature):
        kind = {1: 'bytes', 2: 'float', 3: 'int64'}[f4]
        for f5, w5, value in _fields(lst):
          if f5 != 1:
            continue
          if kind == 'bytes':
            values.append(bytes(value))
          elif kind == 'float':
            if w5 == 2:
              values.extend(struct.unpack('<%df' % (len(value) // 4), value))
            else:
              values.append(struct.unpack('<f', value)[0])
          else:
            if w5 == 2:
              p = 0
              while p < len(value):
                v, p = _varint(value, p)
                values.append(v - (1 << 64) if v >= (1 << 63) else v)
            else:
              values.append(value - (1 << 64) if value >= (1 << 63) else value)
      out[key] = (kind, values)
  return out


# ---- writers (used to build synthetic replay records for tests / benches) -----------------------
def _enc_varint(v):
  out = bytearray()
  v &= (1 << 64) - 1
  while True:
    b = v & 0x7F
    v >>= 7
    if v:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _ld(field, payload):
  return _enc_varint((field << 3) | 2) + _enc_varint(len(payload)) + payload


def make_example(features):
  """features: {key: bytes | list of bytes | list/array of float | list of int}."""
  entries = b''
  for key, value in features.items():
    if isinstance(value, bytes):
      value = [value]
    if len(value) and isinstance(value[0], bytes):
      feature = _ld(1, b''.join(_ld(1, v) for v in value))
    elif len(value) and isinstance(value[0], (int,)) and not isinstance(value[0], bool):
      feature = _ld(3, _ld(1, b''.join(_enc_varint(int(v)) for v in value)))
    else:
      feature = _ld(2, _ld(1, struct.pack('<%df' % len(value), *[float(v) for v in value])))
    entries += _ld(1, _ld(1, key.encode('utf-8')) + _ld(2, feature))
  return _ld(1, entries)


def _feature(value):
  if isinstance(value, bytes):
    value = [value]
  if len(value) and isinstance(value[0], bytes):
    return _ld(1, b''.join(_ld(1, v) for v in value))
  if len(value) and isinstance(value[0], (int,)) and not isinstance(value[0], bool):
    return _ld(3, _ld(1, b''.join(_enc_varint(int(v)) for v in value)))
  return _ld(2, _ld(1, struct.pack('<%df' % len(value), *[float(v) for v in value])))


def make_sequence_example(context, feature_lists):
  """tf.train.SequenceExample{context = 1: Features, feature_lists = 2: FeatureLists{map<string,
  FeatureList{repeated Feature feature = 1}>}} (tensorflow/core/example/example.proto).
  context: {key: value}; feature_lists: {key: [value per step]} with values as in make_example."""
  ctx = b''.join(_ld(1, _ld(1, k.encode('utf-8')) + _ld(2, _feature(v))) for k, v in context.items())
  lists = b''
  for key, steps in feature_lists.items():
    fl = b''.join(_ld(1, _feature(v)) for v in steps)
    lists += _ld(1, _ld(1, key.encode('utf-8')) + _ld(2, fl))
  return _ld(1, ctx) + _ld(2, lists)


def write_tfrecords(path, records):
  with open(path, 'wb') as f:
    for rec in records:
      header = struct.pack('<Q', len(rec))
      f.write(header + struct.pack('<I', masked_crc32c(header)) + rec + struct.pack('<I', masked_crc32c(rec)))
