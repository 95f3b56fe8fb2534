"""The record input path: TFRecord files -> tf.Example parsing -> image decode -> batches.

Replaces the tf.data pipeline of the reference (utils/tfdata.py:64-138 file patterns, :174-210
readers, :273-543 create_parse_tf_example_fn, :629-689 default_input_fn_tmpl) with:

  * TFRecord framing + CRC-32C and the tf.Example wire format walked by the host C++ side of
    libt2r_b200.so (csrc/host_io.cc, no protobuf/TensorFlow), values copied bit-exactly into numpy
    buffers described by a parse plan compiled from the specs;
  * baseline JPEG decode by the split decoder (Huffman on the host worker pool, IDCT / upsampling / colour on the GPU)
    or completely on the host pool (csrc/jpeg_host.cc, csrc/jpeg.cu), both bit-identical with libjpeg-turbo, the
    decoder TensorFlow links; PNG and the JPEG flavours those refuse go through PIL;
  * shuffle / repeat / batch(drop_remainder) with the reference's structure and buffer sizes; up to
    PARSE_PIPELINE_DEPTH batches are parsed / decoded concurrently (the reference's num_parallel_calls).

Semantics kept (SURVEY A-3..A-6): features are looked up by `dataset_key + name` and returned keyed
by spec *path*; specs without a name are not parsed; bfloat16 specs are parsed as float32 and cast;
varlen specs are padded/clipped to shape[0]; encoded images arrive as bytes, '' decodes to zeros, a
decoded size different from the spec raises; only uint8/uint16 image specs are accepted.
"""
import collections
import concurrent.futures
import ctypes as C
import glob
import io
import itertools
import logging
import mmap
import os
import struct

import numpy as np
from PIL import Image
import torch

from tensor2robot_b200 import _lib
from tensor2robot_b200.models import model_interface
from tensor2robot_b200.utils import dtypes
from tensor2robot_b200.utils import tensorspec_utils

ModeKeys = model_interface.ModeKeys
DATA_FORMAT = {'tfrecord': 'tfrecord'}    # recordio / sstable are Google-internal containers
SUPPORTED_PIXEL_ENCODINGS = (dtypes.uint8, dtypes.uint16)
SHUFFLE_BUFFER_SIZE = 500                 # utils/tfdata.py:665
PARSE_PIPELINE_DEPTH = 2                   # batches parsed / decoded concurrently (bounded: each holds its pinned buffers)


def get_batch_size(params, batch_size):
  """params['batch_size'] overrides the configured batch size (utils/tfdata.py:38-61)."""
  params_batch_size = params.get('batch_size') if params else None
  if params_batch_size is None and batch_size is None:
    raise ValueError('Both params["batch_size"] and batch_size are None, one of them has to be set.')
  if params_batch_size is not None:
    return params_batch_size
  return batch_size


def infer_data_format(file_patterns):
  """The container format named inside the pattern string (utils/tfdata.py:64-89)."""
  data_format = None
  for key in ('tfrecord', 'recordio', 'sstable'):
    if key in file_patterns:
      if data_format is not None:
        raise ValueError('More than one data_format {} and {} have been found in {}.'.format(
            key, data_format, file_patterns))
      data_format = key
  if data_format is None:
    raise ValueError('Could not infer file record type from extension of pattern "%s"' % file_patterns)
  return data_format


def get_data_format_and_filenames_list(file_patterns):
  """Comma-separated patterns (optional '<format>:' prefix) -> (format, [files per pattern])."""
  data_format = infer_data_format(file_patterns)
  file_patterns = file_patterns.replace('{}:'.format(data_format), '')
  filenames_list = [sorted(glob.glob(pattern)) for pattern in file_patterns.split(',')]
  for filenames in filenames_list:
    if not filenames:
      raise ValueError('File list for some pattern in {} is empty'.format(file_patterns))
  return data_format, filenames_list


def get_data_format_and_filenames(file_patterns):
  data_format, filenames_list = get_data_format_and_filenames_list(file_patterns)
  return data_format, list(itertools.chain.from_iterable(filenames_list))


# ---------------------------------------------------------------------------------------------
# TFRecord files
# ---------------------------------------------------------------------------------------------
class TFRecordFile(object):
  """A memory-mapped TFRecord file indexed by the C++ reader (CRC-32C verified)."""

  def __init__(self, filename, verify_crc=True):
    self.filename = filename
    size = os.path.getsize(filename)
    self._file = open(filename, 'rb')
    # a private copy-on-write mapping: nothing is read until a record is touched, nothing is ever written back, and
    # (unlike a read-only mapping) ctypes can take its address without copying the shard
    self._map = mmap.mmap(self._file.fileno(), 0, access=mmap.ACCESS_COPY) if size else None
    self._buf = (C.c_char * size).from_buffer(self._map) if size else (C.c_char * 0)()
    self._base = C.addressof(self._buf)
    n = _lib.lib().t2r_tfrecord_index(self._base, size, None, None, 0, 1 if verify_crc else 0)
    if n < 0:
      raise ValueError('%s: %s' % (filename, _lib.last_error()))
    self.offsets = np.zeros(n, np.uint64)
    self.lengths = np.zeros(n, np.uint64)
    if n:
      _lib.lib().t2r_tfrecord_index(self._base, size, self.offsets.ctypes.data, self.lengths.ctypes.data, n, 0)

  def __len__(self):
    return len(self.offsets)

  def pointer(self, i):
    return self._base + int(self.offsets[i]), int(self.lengths[i])

  def record(self, i):
    off, n = int(self.offsets[i]), int(self.lengths[i])
    return bytes(self._buf[off:off + n])

  def __iter__(self):
    for i in range(len(self)):
      yield self.record(i)


def read_records(filename, verify_crc=True):
  """All serialized records of a TFRecord file."""
  return list(TFRecordFile(filename, verify_crc))


# ---------------------------------------------------------------------------------------------
# tf.Example parsing
# ---------------------------------------------------------------------------------------------
def _decode_one(image_bytes, single_img_dims, np_dtype, name):
  if not image_bytes:
    return np.zeros(single_img_dims, np_dtype)          # '' -> zeros (utils/tfdata.py:465-473)
  try:
    with Image.open(io.BytesIO(image_bytes)) as im:
      channels = single_img_dims[2]
      if np_dtype == np.uint16:
        arr = np.asarray(im)
        if arr.dtype != np.uint16:
          arr = arr.astype(np.uint16)
      else:
        im = im.convert('L' if channels == 1 else 'RGB')
        arr = np.asarray(im, dtype=np.uint8)
  except (OSError, SyntaxError, EOFError, struct.error, Image.DecompressionBombError) as e:
    # tf.image.decode_image reports corrupt data as InvalidArgument; PIL raises a family of exception types
    raise ValueError('InvalidArgument: image "%s" cannot be decoded: %s' % (name, e))
  if arr.ndim == 2:
    arr = arr[:, :, None]
  if tuple(arr.shape) != tuple(single_img_dims):
    raise ValueError('InvalidArgument: decoded image "%s" has shape %s, the spec requires %s' % (
        name, tuple(arr.shape), tuple(single_img_dims)))
  return arr


_POOL = None

# Where encoded images are decoded: 'host' (libjpeg-turbo / libpng through PIL on host threads, numpy out) or
# 'device' (baseline JPEGs: Huffman on host threads, IDCT / upsampling / colour on the GPU, uint8 CUDA tensor
# out, bit-identical; anything the split decoder rejects falls back to 'host').  utils/jpeg.py.
# 'auto' (the default) = 'device' on a machine with a GPU - the record path of a training run - else 'host'.
IMAGE_DECODER = os.environ.get('T2R_IMAGE_DECODER', 'auto')


def set_image_decoder(kind):
  global IMAGE_DECODER
  if kind not in ('host', 'device', 'auto'):
    raise ValueError("image decoder must be 'host', 'device' or 'auto'")
  IMAGE_DECODER = kind


def image_decoder():
  """The decoder in effect: 'host' or 'device'."""
  if IMAGE_DECODER != 'auto':
    return IMAGE_DECODER
  import torch
  return 'device' if torch.cuda.is_available() else 'host'


def _pool():
  global _POOL
  if _POOL is None:
    _POOL = concurrent.futures.ThreadPoolExecutor(max_workers=min(32, (os.cpu_count() or 4)))
  return _POOL


def _decode_images(tensor_spec, byte_rows):
  """byte_rows: nested list [B][k] of bytes.  Returns [B, (k,) h, w, c]."""
  if len(tensor_spec.shape) < 3:
    raise ValueError('Shape of tensor spec for image feature "%s" must be 3 dimensional (h, w, c), but is %s' % (
        tensor_spec.name, tensor_spec.shape))
  dims = tuple(tensor_spec.shape[-3:])
  if dims[2] not in (1, 3):
    raise ValueError('Last dimension of shape of tensor spec for image feature "%s" must 1 or 3, but the shape '
                     'is %s' % (tensor_spec.name, tensor_spec.shape))
  if tensor_spec.dtype not in SUPPORTED_PIXEL_ENCODINGS:
    raise ValueError('Decoding an image requires tensorspec.data_type to be uint8 or uint16.')
  np_dtype = tensor_spec.dtype.as_numpy_dtype
  flat = [b for row in byte_rows for b in row]
  per_row = len(byte_rows[0]) if byte_rows else 0
  if image_decoder() == 'device' and np_dtype == np.uint8 and flat and all(b[:2] == b'\xff\xd8' for b in flat):
    from tensor2robot_b200.utils import jpeg
    try:
      out = jpeg.decode_batch(flat, channels=dims[2])
    except jpeg.UnsupportedJpeg:
      out = None
    if out is not None:
      if tuple(out.shape[1:]) != dims:
        raise ValueError('InvalidArgument: decoded image "%s" has shape %s, the spec requires %s' % (
            tensor_spec.name, tuple(out.shape[1:]), dims))
      if len(tensor_spec.shape) > 3 or tensor_spec.varlen_default_value is not None:
        out = out.reshape((len(byte_rows), per_row) + dims)
      return out
  if np_dtype == np.uint8 and flat and all(b[:2] == b'\xff\xd8' for b in flat):
    # baseline JPEGs: the C++ host decoder (csrc/jpeg_host.cc; bit-identical with libjpeg-turbo, real threads - PIL's
    # JPEG plugin decodes under the GIL).  Anything it refuses (progressive, CMYK, another size) goes through PIL below,
    # which decodes it or reports the InvalidArgument.
    from tensor2robot_b200.utils import jpeg
    try:
      out = jpeg.decode_batch_host(flat, dims[0], dims[1], dims[2])
    except jpeg.UnsupportedJpeg:
      out = None
    if out is not None:
      if len(tensor_spec.shape) > 3 or tensor_spec.varlen_default_value is not None:
        out = out.reshape((len(byte_rows), per_row) + dims)
      return out
  if len(flat) >= 8:
    decoded = list(_pool().map(lambda b: _decode_one(b, dims, np_dtype, tensor_spec.name), flat))
  else:
    decoded = [_decode_one(b, dims, np_dtype, tensor_spec.name) for b in flat]
  out = np.stack(decoded) if decoded else np.zeros((0,) + dims, np_dtype)
  per_row = len(byte_rows[0]) if byte_rows else 0
  if len(tensor_spec.shape) > 3 or tensor_spec.varlen_default_value is not None:
    out = out.reshape((len(byte_rows), per_row) + dims)
  return out


def _decode_jpeg_pointers(tensor_spec, addresses, lengths, b, count, owner):
  """The common case of `_decode_images` without materialising the JPEG strings: every image of the batch is present
  and a baseline JPEG that the C++ / split decoder accepts; the decoders read the bytes where the record parser found
  them (inside the mapped record file).  None = take the general path."""
  dims = tuple(tensor_spec.shape[-3:])
  if (len(tensor_spec.shape) < 3 or dims[2] not in (1, 3) or tensor_spec.dtype != dtypes.uint8 or b * count == 0 or
      not addresses.all() or int(lengths.min()) < 4):
    return None
  from tensor2robot_b200.utils import jpeg
  images = jpeg.RecordBytes(addresses, lengths, owner)
  if any(images.head(i, 2) != b'\xff\xd8' for i in range(len(images))):
    return None
  try:
    if image_decoder() == 'device':
      out = jpeg.decode_batch(images, channels=dims[2])
    else:
      out = jpeg.decode_batch_host(images, dims[0], dims[1], dims[2])
  except jpeg.UnsupportedJpeg:
    return None
  if tuple(out.shape[1:]) != dims:
    return None                                   # the general path reports the InvalidArgument
  if len(tensor_spec.shape) > 3 or tensor_spec.varlen_default_value is not None:
    out = out.reshape((b, count) + dims)
  return out


def _records_as_pointers(serialized):
  """list of bytes -> (keep-alive list, pointer array, length array)."""
  n = len(serialized)
  ptrs = (C.c_void_p * n)()
  lens = (C.c_uint64 * n)()
  keep = []
  for i, rec in enumerate(serialized):
    if isinstance(rec, tuple):           # (address, length) from TFRecordFile.pointer
      ptrs[i], lens[i] = rec
    else:
      buf = C.create_string_buffer(rec, len(rec))
      keep.append(buf)
      ptrs[i], lens[i] = C.addressof(buf), len(rec)
  return keep, ptrs, lens


def _parse_sequence_features(ptrs, lens, b, seq_specs, decode_images):
  """tf.io.parse_sequence_example for FixedLenSequenceFeature(allow_missing=True) specs
  (utils/tfdata.py:352-384): {name: spec} -> ({name: [B, T_max, ...]}, {name: int64 [B] lengths})."""
  n = len(seq_specs)
  plans = (_lib.FeaturePlan * n)()
  keep, kinds = [], []
  for i, (key, spec) in enumerate(seq_specs.items()):
    plan = plans[i]
    plan.key = spec.name.encode('utf-8')
    keep.append(plan.key)
    if decode_images and tensorspec_utils.is_encoded_image_spec(spec):
      plan.dtype, plan.count = _lib.T2R_DT_BYTES, (int(spec.shape[0]) if len(spec.shape) > 3 else 1)
      kinds.append('bytes')
    elif spec.dtype == dtypes.string:
      plan.dtype, plan.count = _lib.T2R_DT_BYTES, int(np.prod(tuple(spec.shape))) if spec.shape else 1
      kinds.append('strings')
    elif spec.dtype.is_floating:
      plan.dtype, plan.count = _lib.T2R_DT_FLOAT, int(np.prod(tuple(spec.shape))) if spec.shape else 1
      kinds.append('float')
    elif spec.dtype.is_integer or spec.dtype == dtypes.bool_:
      plan.dtype, plan.count = _lib.T2R_DT_INT64, int(np.prod(tuple(spec.shape))) if spec.shape else 1
      kinds.append('int')
    else:
      raise ValueError('Feature specification with invalid data type for tf.Example parsing: "%s": %s' % (
          key, spec.dtype))
  steps = np.zeros((n, b), np.int64)
  rc = _lib.lib().t2r_sequence_example_parse_batch(ptrs, lens, b, plans, n, 0, steps.ctypes.data)
  if rc != 0:
    raise ValueError('tf.SequenceExample parsing failed: %s' % _lib.last_error())
  t_max = int(steps.max()) if steps.size else 0
  bufs = []
  for i, kind in enumerate(kinds):
    count = plans[i].count
    size = b * max(t_max, 1) * count
    if kind in ('bytes', 'strings'):
      dst, dst_len = np.zeros(size, np.uint64), np.zeros(size, np.uint64)
      plans[i].dst_len = dst_len.ctypes.data
    else:
      dst, dst_len = np.zeros(size, np.float32 if kind == 'float' else np.int64), None
    plans[i].dst = dst.ctypes.data
    bufs.append((dst, dst_len))
  if t_max > 0:
    rc = _lib.lib().t2r_sequence_example_parse_batch(ptrs, lens, b, plans, n, t_max, steps.ctypes.data)
    if rc != 0:
      raise ValueError('tf.SequenceExample parsing failed: %s' % _lib.last_error())
  parsed, lengths = {}, {}
  for i, (key, spec) in enumerate(seq_specs.items()):
    dst, dst_len = bufs[i]
    count, kind = plans[i].count, kinds[i]
    lengths[key] = steps[i].copy()
    if kind in ('bytes', 'strings'):
      rows = [[C.string_at(int(dst[(r * t_max + t) * count + j]), int(dst_len[(r * t_max + t) * count + j]))
               if dst[(r * t_max + t) * count + j] else b'' for t in range(t_max) for j in range(count)]
              for r in range(b)]
      if kind == 'bytes':
        dims = tuple(spec.shape[-3:])
        lead = (t_max, count) if len(spec.shape) > 3 else (t_max,)
        if t_max == 0:
          parsed[key] = np.zeros((b,) + lead + dims, spec.dtype.as_numpy_dtype)
        else:
          step_spec = tensorspec_utils.ExtendedTensorSpec.from_spec(spec, shape=(t_max * count,) + dims)
          parsed[key] = _decode_images(step_spec, rows).reshape((b,) + lead + dims)
      else:
        parsed[key] = np.array(rows, dtype=object).reshape((b, t_max) + tuple(spec.shape))
    else:
      arr = dst[:b * t_max * count].reshape((b, t_max) + tuple(spec.shape))
      if spec.dtype == dtypes.bfloat16:
        import torch
        parsed[key] = torch.from_numpy(arr).to(torch.bfloat16)
      else:
        parsed[key] = arr.astype(spec.dtype.as_numpy_dtype, copy=False)
  del keep
  return parsed, lengths


def _parse_examples(serialized, tensor_spec_dict, decode_images):
  """The tf.parse_example equivalent: {dataset_key+name: spec} -> {same key: numpy [B, ...]}.
  Specs with is_sequence=True switch to tf.io.parse_sequence_example semantics: they are read from the
  SequenceExample's feature_lists, `<name>_length` context specs are not parsed but produced from the
  step counts (utils/tfdata.py:352-384)."""
  b = len(serialized)
  keep, ptrs, lens = _records_as_pointers(serialized)
  seq_specs = collections.OrderedDict((k, v) for k, v in tensor_spec_dict.items() if getattr(v, 'is_sequence', False))
  seq_parsed = {}
  if seq_specs:
    seq_parsed, seq_lengths = _parse_sequence_features(ptrs, lens, b, seq_specs, decode_images)
    for k, v in seq_lengths.items():
      seq_parsed[k + '_length'] = v
    tensor_spec_dict = collections.OrderedDict(
        (k, v) for k, v in tensor_spec_dict.items() if k not in seq_specs and k not in seq_parsed)
    if not tensor_spec_dict:
      del keep
      return seq_parsed
  plans = (_lib.FeaturePlan * len(tensor_spec_dict))()
  buffers = {}
  for i, (key, spec) in enumerate(tensor_spec_dict.items()):
    plan = plans[i]
    plan.key = spec.name.encode('utf-8')
    keep.append(plan.key)
    encoded = decode_images and tensorspec_utils.is_encoded_image_spec(spec)
    varlen = spec.varlen_default_value is not None
    if encoded:
      count = int(spec.shape[0]) if (len(spec.shape) > 3 or varlen) else 1
      dst = np.zeros(b * count, np.uint64)
      dst_len = np.zeros(b * count, np.uint64)
      plan.dtype, plan.count = _lib.T2R_DT_BYTES, (-count if varlen else count)
      buffers[key] = ('bytes', dst, dst_len, count)
    else:
      shape = tuple(spec.shape)
      count = int(np.prod(shape)) if shape else 1
      if varlen:
        count = int(spec.shape[0])
      if spec.dtype == dtypes.string:
        dst, dst_len = np.zeros(b * count, np.uint64), np.zeros(b * count, np.uint64)
        plan.dtype = _lib.T2R_DT_BYTES
        buffers[key] = ('strings', dst, dst_len, count)
      elif spec.dtype.is_floating:
        dst, dst_len = np.zeros(b * count, np.float32), np.zeros(b, np.uint64)
        plan.dtype = _lib.T2R_DT_FLOAT
        plan.pad_float = float(spec.varlen_default_value) if varlen else 0.0
        buffers[key] = ('float', dst, dst_len, count)
      elif spec.dtype.is_integer or spec.dtype == dtypes.bool_:
        dst, dst_len = np.zeros(b * count, np.int64), np.zeros(b, np.uint64)
        plan.dtype = _lib.T2R_DT_INT64
        plan.pad_int64 = int(spec.varlen_default_value) if varlen else 0
        buffers[key] = ('int', dst, dst_len, count)
      else:
        raise ValueError('unsupported dtype %s for feature %s' % (spec.dtype, spec.name))
      plan.count = -count if varlen else count
    plan.required = 0 if (getattr(spec, 'is_optional', False) or varlen) else 1
    plan.dst = dst.ctypes.data
    plan.dst_len = dst_len.ctypes.data
  rc = _lib.lib().t2r_example_parse_batch(ptrs, lens, b, plans, len(tensor_spec_dict))
  if rc != 0:
    raise ValueError('tf.Example parsing failed: %s' % _lib.last_error())
  parsed = {}
  for key, spec in tensor_spec_dict.items():
    kind, dst, dst_len, count = buffers[key]
    if kind in ('bytes', 'strings'):
      if kind == 'bytes':
        fast = _decode_jpeg_pointers(spec, dst, dst_len, b, count, keep)
        if fast is not None:
          parsed[key] = fast
          continue
      rows = [[C.string_at(int(dst[r * count + j]), int(dst_len[r * count + j])) if dst[r * count + j] else b''
               for j in range(count)] for r in range(b)]
      if kind == 'bytes':
        parsed[key] = _decode_images(spec, rows)
      else:
        arr = np.array(rows, dtype=object)
        parsed[key] = arr.reshape((b,) + tuple(spec.shape)) if spec.shape else arr.reshape(b)
    else:
      shape = (int(spec.shape[0]),) if spec.varlen_default_value is not None else tuple(spec.shape)
      arr = dst.reshape((b,) + shape)
      if spec.dtype == dtypes.bfloat16:
        import torch
        parsed[key] = torch.from_numpy(arr).to(torch.bfloat16)      # parsed as f32 then cast (:326-346)
      else:
        parsed[key] = arr.astype(spec.dtype.as_numpy_dtype, copy=False)
  del keep
  parsed.update(seq_parsed)
  return parsed


def create_parse_tf_example_fn(feature_tspec, label_tspec=None, decode_images=True):
  """Returns parse_tf_example_fn(serialized) -> features | (features, labels).

  `serialized` is a list of serialized tf.Example protos (bytes, or (address, length) pairs from
  TFRecordFile.pointer) or a {dataset_key: list} dict when specs carry dataset keys."""

  def parse_tf_example_fn(*input_values):
    serialized = input_values[-1]
    if not isinstance(serialized, dict):
      serialized = {'': serialized}
    parsed_tensors = {}
    for dataset_key, records in serialized.items():
      spec_dict = {}
      for tspec in (feature_tspec, label_tspec):
        if tspec is None:
          continue
        subset = tensorspec_utils.filter_spec_structure_by_dataset(tspec, dataset_key)
        _, tensor_spec_dict = tensorspec_utils.tensorspec_to_feature_dict(subset, decode_images=decode_images)
        for name, spec in tensor_spec_dict.items():
          if name in spec_dict:
            tensorspec_utils.assert_equal_spec_or_tensor(spec_dict[name], spec)
          spec_dict[name] = spec
      for name, value in _parse_examples(records, spec_dict, decode_images).items():
        parsed_tensors[dataset_key + name] = value

    def pack(tspec):
      flat = tensorspec_utils.TensorSpecStruct(sorted(tensorspec_utils.flatten_spec_structure(tspec).items()))
      out = tensorspec_utils.TensorSpecStruct()
      for key, value in flat.items():
        if value.name is None:
          continue
        lookup = value.dataset_key + value.name
        if lookup in parsed_tensors:
          out[key] = parsed_tensors[lookup]
      return tensorspec_utils.validate_and_pack(flat, out, ignore_batch=True)

    features = pack(feature_tspec)
    if label_tspec is not None:
      return features, pack(label_tspec)
    return features

  return parse_tf_example_fn


# ---------------------------------------------------------------------------------------------
# the input pipeline
# ---------------------------------------------------------------------------------------------
def record_stream(filenames, mode, seed=None, shard=(0, 1), verify_crc=True, keepalive=None):
  """Yields (address, length) record pointers: files shuffled and repeated forever in TRAIN, a single
  ordered pass otherwise; `shard=(rank, world)` keeps every world-th file (or record when there are
  fewer files than ranks).  The pointers are into the files' mappings: a consumer that uses them after this generator
  has finished passes a `keepalive` list, which receives every opened file."""
  rank, world = shard
  rng = np.random.RandomState(seed)
  files = list(filenames)
  by_record = len(files) < world
  if not by_record:
    files = files[rank::world]
  opened = {}
  while True:
    order = list(range(len(files)))
    if mode == ModeKeys.TRAIN:
      rng.shuffle(order)
    for fi in order:
      f = opened.get(fi)
      if f is None:
        f = opened[fi] = TFRecordFile(files[fi], verify_crc)
        if keepalive is not None:
          keepalive.append(f)
      idx = range(rank, len(f), world) if by_record else range(len(f))
      for i in idx:
        yield f.pointer(i)
    if mode != ModeKeys.TRAIN:
      return


def shuffled(stream, buffer_size, seed=None):
  """tf.data shuffle(buffer_size)."""
  rng = np.random.RandomState(seed)
  buf = []
  for item in stream:
    if len(buf) < buffer_size:
      buf.append(item)
      continue
    j = rng.randint(0, buffer_size)
    yield buf[j]
    buf[j] = item
  rng.shuffle(buf)
  for item in buf:
    yield item


def default_input_fn_tmpl(file_patterns, batch_size, feature_spec, label_spec, num_parallel_calls=4,
                          is_training=False, preprocess_fn=None, shuffle_filenames=True,
                          shuffle_buffer_size=SHUFFLE_BUFFER_SIZE, mode=ModeKeys.TRAIN, seed=None, shard=(0, 1),
                          **unused):
  """Generator of (features, labels) batches: list files -> shuffle -> repeat -> batch(drop_remainder)
  -> parse -> preprocess (utils/tfdata.py:629-689)."""
  del shuffle_filenames, unused
  mapped = []        # the opened record files: batches in flight hold pointers into their mappings
  if isinstance(file_patterns, dict):
    streams = {}
    for key, patterns in file_patterns.items():
      _, filenames = get_data_format_and_filenames(patterns)
      streams[key] = record_stream(filenames, mode, seed, shard, keepalive=mapped)
  else:
    _, filenames = get_data_format_and_filenames(file_patterns)
    streams = {'': record_stream(filenames, mode, seed, shard, keepalive=mapped)}
  if is_training or mode == ModeKeys.TRAIN:
    streams = {k: shuffled(s, shuffle_buffer_size, seed) for k, s in streams.items()}
  parse_fn = create_parse_tf_example_fn(feature_spec, label_spec)

  def draw():
    batch = {k: list(itertools.islice(s, batch_size)) for k, s in streams.items()}
    if any(len(v) < batch_size for v in batch.values()):
      return None                                # drop_remainder=True
    return batch if len(batch) > 1 or '' not in batch else batch['']

  def finish(parsed):
    features, labels = parsed if label_spec is not None else (parsed, None)
    if preprocess_fn is not None:
      features, labels = preprocess_fn(features, labels)
    return features, labels

  in_flight = max(1, min(int(num_parallel_calls or 1), PARSE_PIPELINE_DEPTH))
  if in_flight == 1:
    while True:
      batch = draw()
      if batch is None:
        return
      yield finish(parse_fn(batch))
  # The reference's `num_parallel_calls` (parallel map over parse): up to PARSE_PIPELINE_DEPTH batches are parsed
  # concurrently and yielded in order.  The serial parts of one batch (CRC, wire parse, copies of the image strings) then
  # overlap the threaded Huffman stage of its neighbour; records are still drawn in order on this thread.
  cuda = torch.cuda.is_available() and torch.cuda.is_initialized()
  device = torch.cuda.current_device() if cuda else None

  def work(batch, stream):
    if stream is None:
      return parse_fn(batch)
    torch.cuda.set_device(device)                # device and stream are per-thread state
    with torch.cuda.stream(stream):
      return parse_fn(batch)

  pool = concurrent.futures.ThreadPoolExecutor(max_workers=in_flight, thread_name_prefix='t2r-parse')
  pending = collections.deque()
  try:
    exhausted = False
    while True:
      while not exhausted and len(pending) < in_flight:
        batch = draw()
        if batch is None:
          exhausted = True
          break
        # the device half of the split JPEG decoder launches on the stream that is current HERE at submit time
        stream = torch.cuda.current_stream() if cuda else None
        pending.append(pool.submit(work, batch, stream))
      if not pending:
        return
      yield finish(pending.popleft().result())
  finally:
    for f in pending:
      f.cancel()
    pool.shutdown(wait=True)       # a parse that is still running reads the mappings in `mapped`
    del mapped


def get_input_fn(feature_spec, label_spec, file_patterns, mode, batch_size, preprocess_fn=None, **kwargs):
  """Returns input_fn(params) -> iterator of (features, labels) (utils/tfdata.py:692-718)."""

  def input_fn(params=None):
    return default_input_fn_tmpl(file_patterns=file_patterns, batch_size=get_batch_size(params, batch_size),
                                 feature_spec=feature_spec, label_spec=label_spec,
                                 is_training=(mode == ModeKeys.TRAIN), preprocess_fn=preprocess_fn, mode=mode,
                                 **kwargs)
  return input_fn


# ---------------------------------------------------------------------------------------------
# Remaining helpers of the reference module (utils/tfdata.py:143-238, 546-626), as host-side iterators.
# ---------------------------------------------------------------------------------------------
def get_dataset_metadata(file_patterns):
  """(data_format, num_shards, approximate num_examples_per_shard) for shuffling parameters (:143-171): the first
  shard is indexed; like the reference the count starts at one ("at least one example per shard")."""
  data_format, files = get_data_format_and_filenames(file_patterns)
  logging.info('Estimating dataset size from %s...', files[0])
  return data_format, len(files), 1 + len(TFRecordFile(files[0]))


def parallel_read(file_patterns, num_shards=None, num_readers=None, num_epochs=None, seed=None):
  """Serialized tf.Example protos of the files (:174-210): shards shuffled per epoch and read `num_readers` at a time,
  one record from each in turn (the interleave that breaks intra-shard correlation); repeats `num_epochs` times,
  forever if None."""
  _, filenames = get_data_format_and_filenames(file_patterns)
  if num_shards is None:
    num_shards = len(filenames)
  if num_readers is None:
    num_readers = num_shards
  rng = np.random.RandomState(seed)
  opened = {}
  epoch = 0
  while num_epochs is None or epoch < num_epochs:
    epoch += 1
    order = list(rng.permutation(len(filenames)))
    active = []
    while order or active:
      while order and len(active) < max(1, num_readers):
        name = filenames[order.pop(0)]
        if name not in opened:
          opened[name] = TFRecordFile(name)
        active.append(iter(opened[name]))
      for reader in list(active):
        try:
          yield next(reader)
        except StopIteration:
          active.remove(reader)


def serialized_to_parsed(dataset, feature_tspec, label_tspec, num_parallel_calls=2):
  """Maps the spec-generated parser over an iterable of serialized batches - lists of protos, or {dataset_key: list}
  dicts (:213-238).  num_parallel_calls is accepted for signature parity; the C++ parser is already batched."""
  del num_parallel_calls
  parse_tf_example_fn = create_parse_tf_example_fn(feature_tspec=feature_tspec, label_tspec=label_tspec)
  for serialized in dataset:
    yield parse_tf_example_fn(serialized)


def _jpeg_specs(spec):
  return [(key, value) for key, value in tensorspec_utils.flatten_spec_structure(spec).items()
          if getattr(value, 'data_format', None) == 'jpeg'] if spec is not None else []


def create_compress_fn(feature_spec, label_spec, quality=90):
  """compress_fn(features, labels=None): every tensor whose spec has data_format 'jpeg' becomes an object array of
  JPEG strings, one per batch element (:546-585; the infeed compression of the TPU path).  Host numpy in, host out."""

  def compress(tensor):
    tensor = np.asarray(tensor)
    if tensor.dtype != np.uint8:       # tf.image.convert_image_dtype(float -> uint8): scale, round half up, saturate
      tensor = np.clip(np.floor(tensor.astype(np.float32) * 255.0 + 0.5), 0, 255).astype(np.uint8)
    out = np.empty(tensor.shape[0], dtype=object)
    for i, img in enumerate(tensor):
      buf = io.BytesIO()
      Image.fromarray(img[..., 0] if img.shape[-1] == 1 else img).save(buf, format='JPEG', quality=quality)
      out[i] = buf.getvalue()
    return out

  def compress_fn(features, labels=None):
    for key, _ in _jpeg_specs(feature_spec):
      features[key] = compress(features[key])
    if labels is not None:
      for key, _ in _jpeg_specs(label_spec):
        labels[key] = compress(labels[key])
    return features, labels

  return compress_fn


def create_decompress_fn(feature_spec, label_spec):
  """The inverse of create_compress_fn (:588-626): JPEG strings -> [batch] + spec.shape arrays in the spec's dtype
  (floats in [0, 1])."""

  def decompress(strings, spec):
    shape = tuple(spec.shape)
    images = np.stack([np.asarray(Image.open(io.BytesIO(s))).reshape(shape) for s in strings])
    np_dtype = spec.dtype.as_numpy_dtype
    if np.issubdtype(np_dtype, np.floating):
      return (images.astype(np.float32) / np.float32(255.0)).astype(np_dtype)
    return images.astype(np_dtype)

  def decompress_fn(features, labels=None):
    for key, value in _jpeg_specs(feature_spec):
      features[key] = decompress(features[key], value)
    if labels is not None:
      for key, value in _jpeg_specs(label_spec):
        labels[key] = decompress(labels[key], value)
    return features, labels

  return decompress_fn
