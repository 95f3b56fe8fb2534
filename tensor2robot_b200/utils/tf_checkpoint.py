"""Reader of TensorFlow checkpoints in the tensor-bundle format (`<prefix>.index` + `<prefix>.data-NNNNN-of-MMMMM`), so
that weights trained with the reference can be loaded by variable name (models/abstract_model.py:87-126:
`default_init_from_checkpoint_fn` -> tf.train.load_checkpoint / init_from_checkpoint).

TensorFlow is not available here, so the published on-disk format is read directly:
* the index is a LevelDB-style sorted string table (tensorflow/core/lib/io/table*.cc): prefix-compressed key / value
  blocks, each followed by a 1-byte compression type (0 none, 1 snappy) and a masked CRC-32C; a 48-byte footer holds the
  handles of the meta-index and index blocks and the magic 0xdb4775248b80fb57;
* key "" maps to a BundleHeaderProto, every other key (a variable name) to a BundleEntryProto {dtype = 1, shape = 2,
  shard_id = 3, offset = 4, size = 5, crc32c = 6 (masked CRC-32C of the tensor bytes)}
  (tensorflow/core/protobuf/tensor_bundle.proto);
* tensor bytes are raw little-endian arrays at [offset, offset + size) of the shard file.
Pinned by the reference's own fixture test_data/mock_exported_savedmodel/variables (tests/test_tf_checkpoint.py): every
block and tensor checksum stored by TensorFlow verifies."""
import collections
import ctypes as C
import os
import struct

import numpy as np

from tensor2robot_b200 import _lib

_TABLE_MAGIC = 0xdb4775248b80fb57
_FOOTER_LEN = 48
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.dtype('<f4'), 2: np.dtype('<f8'), 3: np.dtype('<i4'), 4: np.dtype('u1'), 5: np.dtype('<i2'),
           6: np.dtype('i1'), 9: np.dtype('<i8'), 10: np.dtype('?'), 17: np.dtype('<u2'), 19: np.dtype('<f2'),
           22: np.dtype('<u4'), 23: np.dtype('<u8')}
_DT_BFLOAT16 = 14

BundleEntry = collections.namedtuple('BundleEntry', ['dtype', 'shape', 'shard_id', 'offset', 'size', 'crc32c'])


class CheckpointError(ValueError):
  pass


def _crc32c(data):
  buf = (C.c_uint8 * len(data)).from_buffer_copy(data) if data else None
  return int(_lib.lib().t2r_crc32c(buf, len(data))) & 0xFFFFFFFF


def _mask(crc):
  return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF


def _varint(buf, pos):
  result = shift = 0
  while True:
    if pos >= len(buf):
      raise CheckpointError('truncated varint')
    b = buf[pos]
    pos += 1
    result |= (b & 0x7F) << shift
    if not b & 0x80:
      return result, pos
    shift += 7
    if shift > 63:
      raise CheckpointError('varint too long')


def snappy_decompress(data):
  """The raw Snappy block format: varint uncompressed length, then literals (tag & 3 == 0) and back-references with
  1-, 2- or 4-byte offsets (copies may overlap their own output)."""
  total, pos = _varint(data, 0)
  out = bytearray()
  n = len(data)
  while pos < n:
    tag = data[pos]
    pos += 1
    kind = tag & 3
    if kind == 0:
      length = tag >> 2
      if length >= 60:
        extra = length - 59
        length = int.from_bytes(data[pos:pos + extra], 'little')
        pos += extra
      length += 1
      if pos + length > n:
        raise CheckpointError('snappy literal overruns the input')
      out += data[pos:pos + length]
      pos += length
      continue
    if kind == 1:
      length = ((tag >> 2) & 7) + 4
      offset = ((tag >> 5) << 8) | data[pos]
      pos += 1
    elif kind == 2:
      length = (tag >> 2) + 1
      offset = data[pos] | (data[pos + 1] << 8)
      pos += 2
    else:
      length = (tag >> 2) + 1
      offset = int.from_bytes(data[pos:pos + 4], 'little')
      pos += 4
    if offset == 0 or offset > len(out):
      raise CheckpointError('snappy back-reference outside the output')
    start = len(out) - offset
    for i in range(length):           # byte-wise: the source may overlap the bytes being written
      out.append(out[start + i])
  if len(out) != total:
    raise CheckpointError('snappy stream decodes to %d bytes, header says %d' % (len(out), total))
  return bytes(out)


def _read_block(data, offset, size, verify):
  """Block contents at (offset, size) + the 5-byte trailer {type, masked crc32c(contents + type)}."""
  end = offset + size
  if end + 5 > len(data):
    raise CheckpointError('block handle outside the file')
  contents, kind = data[offset:end], data[end]
  (stored,) = struct.unpack_from('<I', data, end + 1)
  if verify and _mask(_crc32c(data[offset:end + 1])) != stored:
    raise CheckpointError('block checksum mismatch at offset %d' % offset)
  if kind == 0:
    return contents
  if kind == 1:
    return snappy_decompress(contents)
  raise CheckpointError('unknown block compression type %d' % kind)


def _iter_block(block):
  """(key, value) pairs of a table block: entries {shared, non_shared, value_len varints, key suffix, value}, then the
  restart array and its uint32 count."""
  (num_restarts,) = struct.unpack_from('<I', block, len(block) - 4)
  limit = len(block) - 4 - 4 * num_restarts
  pos, key = 0, b''
  while pos < limit:
    shared, pos = _varint(block, pos)
    non_shared, pos = _varint(block, pos)
    value_len, pos = _varint(block, pos)
    if shared > len(key):
      raise CheckpointError('corrupt key prefix')
    key = key[:shared] + block[pos:pos + non_shared]
    pos += non_shared
    yield key, block[pos:pos + value_len]
    pos += value_len


def _block_handle(buf, pos=0):
  offset, pos = _varint(buf, pos)
  size, pos = _varint(buf, pos)
  return offset, size, pos


def _fields(buf):
  """(field number, wire type, value) triples of a protobuf message."""
  pos = 0
  while pos < len(buf):
    tag, pos = _varint(buf, pos)
    field, wire = tag >> 3, tag & 7
    if wire == 0:
      value, pos = _varint(buf, pos)
    elif wire == 1:
      value = buf[pos:pos + 8]
      pos += 8
    elif wire == 2:
      length, pos = _varint(buf, pos)
      value = buf[pos:pos + length]
      pos += length
    elif wire == 5:
      value = buf[pos:pos + 4]
      pos += 4
    else:
      raise CheckpointError('unsupported protobuf wire type %d' % wire)
    yield field, wire, value


def _signed(v):
  return v - (1 << 64) if v >= (1 << 63) else v


def _expect(wire, wanted, what):
  if wire != wanted:
    raise CheckpointError('%s has protobuf wire type %d, expected %d' % (what, wire, wanted))


def _parse_entry(buf):
  dtype, shape, shard_id, offset, size, crc = 0, [], 0, 0, 0, 0
  for field, wire, value in _fields(buf):
    if field == 1:
      _expect(wire, 0, 'BundleEntryProto.dtype')
      dtype = value
    elif field == 2:                      # TensorShapeProto{repeated Dim dim = 2 {int64 size = 1}}
      _expect(wire, 2, 'BundleEntryProto.shape')
      for f2, w2, dim in _fields(value):
        if f2 == 2:
          _expect(w2, 2, 'TensorShapeProto.dim')
          dim_size = 0
          for f3, w3, v3 in _fields(dim):
            if f3 == 1:
              _expect(w3, 0, 'TensorShapeProto.Dim.size')
              dim_size = _signed(v3)
          shape.append(dim_size)
    elif field == 3:
      _expect(wire, 0, 'BundleEntryProto.shard_id')
      shard_id = value
    elif field == 4:
      _expect(wire, 0, 'BundleEntryProto.offset')
      offset = value
    elif field == 5:
      _expect(wire, 0, 'BundleEntryProto.size')
      size = value
    elif field == 6:
      _expect(wire, 5, 'BundleEntryProto.crc32c')
      (crc,) = struct.unpack('<I', value)
    elif field == 7:
      raise CheckpointError('partitioned (sliced) variables are not supported')
  if any(d < 0 for d in shape):
    raise CheckpointError('negative tensor dimension')
  return BundleEntry(dtype, tuple(shape), shard_id, offset, size, crc)


class CheckpointReader(object):
  """tf.train.load_checkpoint(prefix): has_tensor / get_tensor / get_variable_to_shape_map / ..._dtype_map."""

  def __init__(self, prefix, verify=True):
    self._prefix = prefix
    self._verify = verify
    self._entries = collections.OrderedDict()
    self.num_shards = 1
    self._shards = {}
    with open(prefix + '.index', 'rb') as f:
      data = f.read()
    try:
      self._parse_index(data)
    except CheckpointError:
      raise
    except (IndexError, struct.error, UnicodeDecodeError, OverflowError, ValueError, TypeError) as e:
      raise CheckpointError('%s.index is corrupt: %s' % (prefix, e))

  def _parse_index(self, data):
    prefix, verify = self._prefix, self._verify
    if len(data) < _FOOTER_LEN or struct.unpack_from('<Q', data, len(data) - 8)[0] != _TABLE_MAGIC:
      raise CheckpointError('%s.index is not a TensorFlow tensor-bundle index (bad table magic)' % prefix)
    footer = data[-_FOOTER_LEN:]
    _, _, pos = _block_handle(footer)               # meta-index block: unused by bundles
    index_offset, index_size, _ = _block_handle(footer, pos)
    for _, handle in _iter_block(_read_block(data, index_offset, index_size, verify)):
      offset, size, _ = _block_handle(handle)
      for key, value in _iter_block(_read_block(data, offset, size, verify)):
        if key == b'':
          for field, _w, v in _fields(value):         # BundleHeaderProto{num_shards = 1, endianness = 2}
            if field == 1:
              _expect(_w, 0, 'BundleHeaderProto.num_shards')
              self.num_shards = v
            elif field == 2 and v != 0:
              raise CheckpointError('big-endian bundles are not supported')
        else:
          self._entries[key.decode('utf-8')] = _parse_entry(value)

  def _shard(self, shard_id):
    if shard_id not in self._shards:
      path = '%s.data-%05d-of-%05d' % (self._prefix, shard_id, self.num_shards)
      if not os.path.exists(path):
        raise CheckpointError('data shard %s is missing' % path)
      self._shards[shard_id] = np.memmap(path, dtype=np.uint8, mode='r') if os.path.getsize(path) else np.zeros(0, np.uint8)
    return self._shards[shard_id]

  def has_tensor(self, name):
    return name in self._entries

  def get_variable_to_shape_map(self):
    return {k: list(e.shape) for k, e in self._entries.items()}

  def get_variable_to_dtype_map(self):
    return {k: ('bfloat16' if e.dtype == _DT_BFLOAT16 else _DTYPES[e.dtype].name) for k, e in self._entries.items()
            if e.dtype == _DT_BFLOAT16 or e.dtype in _DTYPES}

  def get_tensor(self, name):
    """numpy array of the variable (bfloat16 is widened to float32)."""
    if name not in self._entries:
      raise KeyError('tensor %s is not in checkpoint %s' % (name, self._prefix))
    e = self._entries[name]
    raw = bytes(self._shard(e.shard_id)[e.offset:e.offset + e.size])
    if len(raw) != e.size:
      raise CheckpointError('tensor %s lies outside its data shard' % name)
    if self._verify and _mask(_crc32c(raw)) != e.crc32c:
      raise CheckpointError('tensor %s: checksum mismatch' % name)
    if e.dtype == _DT_BFLOAT16:
      wide = np.frombuffer(raw, dtype='<u2').astype(np.uint32) << 16
      return wide.view(np.float32).reshape(e.shape)
    if e.dtype not in _DTYPES:
      raise CheckpointError('tensor %s has unsupported dtype enum %d' % (name, e.dtype))
    dt = _DTYPES[e.dtype]
    count = int(np.prod(e.shape)) if e.shape else 1
    if count * dt.itemsize != e.size:
      raise CheckpointError('tensor %s: %d bytes for shape %s of %s' % (name, e.size, e.shape, dt.name))
    return np.frombuffer(raw, dtype=dt).reshape(e.shape).copy()


def load_checkpoint(ckpt_dir_or_file):
  """tf.train.load_checkpoint: a prefix, or a directory holding a `checkpoint` state file / a single bundle."""
  prefix = ckpt_dir_or_file
  if os.path.isdir(prefix):
    state = os.path.join(prefix, 'checkpoint')
    found = None
    if os.path.exists(state):
      with open(state) as f:
        for line in f:
          if line.startswith('model_checkpoint_path:'):
            found = line.split(':', 1)[1].strip().strip('"')
      if found and not os.path.isabs(found):
        found = os.path.join(prefix, found)
    if not found:
      indexes = sorted(p for p in os.listdir(prefix) if p.endswith('.index'))
      if len(indexes) != 1:
        raise CheckpointError('cannot pick a checkpoint in %s' % prefix)
      found = os.path.join(prefix, indexes[0][:-len('.index')])
    prefix = found
  return CheckpointReader(prefix)


# ---------------------------------------------------------------------------------------------
# Writer: the same format, uncompressed blocks, one data shard.
# ---------------------------------------------------------------------------------------------
_NP_TO_ENUM = {np.dtype(v).newbyteorder('='): k for k, v in _DTYPES.items()}
_BLOCK_TARGET = 4096


def _enc_varint(value):
  value &= (1 << 64) - 1
  out = bytearray()
  while True:
    b = value & 0x7F
    value >>= 7
    if value:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _encode_entry(dtype_enum, shape, offset, size, crc):
  """BundleEntryProto in TensorFlow's field order; proto3 defaults (shard_id 0, offset 0) are omitted."""
  dims = b''.join(b'\x12' + _enc_varint(len(d)) + d for d in (b'\x08' + _enc_varint(s) for s in shape))
  out = b'\x08' + _enc_varint(dtype_enum) + b'\x12' + _enc_varint(len(dims)) + dims
  if offset:
    out += b'\x20' + _enc_varint(offset)
  if size:
    out += b'\x28' + _enc_varint(size)
  return out + b'\x35' + struct.pack('<I', crc)


def _build_block(items):
  """Table block with a restart point at every entry (no key prefix sharing)."""
  body, restarts = bytearray(), []
  for key, value in items:
    restarts.append(len(body))
    body += _enc_varint(0) + _enc_varint(len(key)) + _enc_varint(len(value)) + key + value
  if not restarts:
    restarts = [0]
  body += b''.join(struct.pack('<I', r) for r in restarts) + struct.pack('<I', len(restarts))
  return bytes(body)


def write_checkpoint(prefix, tensors):
  """Writes {variable name: numpy array} as `<prefix>.index` + `<prefix>.data-00000-of-00001`, readable by
  tf.train.load_checkpoint / this module's CheckpointReader.  Tensors are laid out in key order like BundleWriter does."""
  names = sorted(tensors, key=lambda n: n.encode('utf-8'))
  dirname = os.path.dirname(prefix)
  if dirname:
    os.makedirs(dirname, exist_ok=True)
  header = b'\x08\x01' + b'\x1a\x02\x08\x01'          # BundleHeaderProto{num_shards: 1, version{producer: 1}}
  items, offset = [(b'', header)], 0
  with open(prefix + '.data-00000-of-00001', 'wb') as data_file:
    for name in names:
      if not name:
        raise CheckpointError('empty variable name')
      a = np.asarray(tensors[name])
      dt = a.dtype.newbyteorder('=') if a.dtype.byteorder in '<>' else a.dtype
      if dt not in _NP_TO_ENUM:
        raise CheckpointError('variable %s: dtype %s cannot be stored' % (name, a.dtype))
      raw = np.ascontiguousarray(a, dtype=np.dtype(dt).newbyteorder('<')).tobytes()
      data_file.write(raw)
      items.append((name.encode('utf-8'), _encode_entry(_NP_TO_ENUM[dt], a.shape, offset, len(raw), _mask(_crc32c(raw)))))
      offset += len(raw)
  out, index_items, block, block_bytes = bytearray(), [], [], 0

  def emit(contents):
    handle = _enc_varint(len(out)) + _enc_varint(len(contents))
    out.extend(contents + b'\x00' + struct.pack('<I', _mask(_crc32c(contents + b'\x00'))))
    return handle

  for key, value in items:
    block.append((key, value))
    block_bytes += len(key) + len(value) + 3
    if block_bytes >= _BLOCK_TARGET:
      index_items.append((key, emit(_build_block(block))))
      block, block_bytes = [], 0
  if block:
    index_items.append((block[-1][0], emit(_build_block(block))))
  meta_handle = emit(_build_block([]))
  index_handle = emit(_build_block(index_items))
  footer = meta_handle + index_handle
  out.extend(footer + b'\x00' * (_FOOTER_LEN - 8 - len(footer)) + struct.pack('<Q', _TABLE_MAGIC))
  with open(prefix + '.index', 'wb') as f:
    f.write(bytes(out))


def raw_index_entries(prefix):
  """{key bytes: value bytes} of an index file (decompressed): used by the tests to compare encodings."""
  with open(prefix + '.index', 'rb') as f:
    data = f.read()
  footer = data[-_FOOTER_LEN:]
  _, _, pos = _block_handle(footer)
  index_offset, index_size, _ = _block_handle(footer, pos)
  entries = collections.OrderedDict()
  for _, handle in _iter_block(_read_block(data, index_offset, index_size, True)):
    offset, size, _ = _block_handle(handle)
    for key, value in _iter_block(_read_block(data, offset, size, True)):
      entries[key] = value
  return entries
