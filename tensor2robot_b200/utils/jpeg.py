"""Split JPEG decoding: Huffman entropy decoding on host threads (csrc/jpeg_host.cc), inverse DCT /
chroma upsampling / colour conversion on the GPU (csrc/jpeg.cu).  Bit-exact with libjpeg(-turbo)'s
defaults, i.e. with the `tf.image.decode_image` call of the reference's parser (utils/tfdata.py:426-484).

Supported: baseline sequential Huffman JPEG, 8 bit, one interleaved scan, restart intervals, greyscale or
YCbCr 4:4:4 / 4:2:2 / 4:2:0.  Everything else raises `UnsupportedJpeg` (callers keep the host decoder for
those)."""
import ctypes as C

import numpy as np
import torch

from tensor2robot_b200 import _lib


class UnsupportedJpeg(ValueError):
  pass


def _address_of(data, keep):
  """Address of the bytes of `data` for a read-only C call; `keep` holds whatever must outlive the call.  A `bytes`
  object is read in place (no copy of a 100 KB JPEG string per image and batch); anything else is copied."""
  if isinstance(data, bytes):
    keep.append(data)
    return C.cast(C.c_char_p(data), C.c_void_p).value or 0
  buf = (C.c_char * len(data)).from_buffer_copy(data)
  keep.append(buf)
  return C.addressof(buf)


def parse(data):
  """Header fields of one JPEG stream as a `_lib.JpegInfo`."""
  info = _lib.JpegInfo()
  keep = []
  rc = _lib.lib().t2r_jpeg_parse(C.c_void_p(_address_of(data, keep)), len(data), C.byref(info))
  if rc != 0:
    raise UnsupportedJpeg(_lib.last_error())
  return info


class RecordBytes(object):
  """A batch of byte strings that live in somebody else's memory (tf.Example payloads inside a mapped record file):
  `addresses` / `lengths` are uint64 arrays, `owner` whatever keeps that memory valid.  The decoders read them in place."""

  def __init__(self, addresses, lengths, owner=None):
    self.addresses = np.ascontiguousarray(addresses, np.uint64)
    self.lengths = np.ascontiguousarray(lengths, np.uint64)
    self.owner = owner

  def __len__(self):
    return int(self.addresses.shape[0])

  def head(self, i, n):
    return C.string_at(int(self.addresses[i]), min(n, int(self.lengths[i])))


def _pointer_arrays(images, keep):
  """(void* array, uint64 array) over a list of bytes or a RecordBytes."""
  b = len(images)
  ptrs = (C.c_void_p * b)()
  lens = (C.c_uint64 * b)()
  if isinstance(images, RecordBytes):
    C.memmove(ptrs, images.addresses.ctypes.data, 8 * b)
    C.memmove(lens, images.lengths.ctypes.data, 8 * b)
    keep.append(images)
  else:
    for i, img in enumerate(images):
      ptrs[i], lens[i] = _address_of(img, keep), len(img)
  return ptrs, lens


_INFO_DTYPE = np.dtype([('struct_size', '<u4'), ('width', '<i4'), ('height', '<i4'), ('ncomp', '<i4'),
                        ('comp_id', '<i4', 3), ('h', '<i4', 3), ('v', '<i4', 3), ('tq', '<i4', 3), ('hmax', '<i4'),
                        ('vmax', '<i4'), ('mcux', '<i4'), ('mcuy', '<i4'), ('restart_interval', '<i4'),
                        ('reserved', '<i4'), ('coef_offset', '<i8', 3), ('coef_count', '<i8'), ('qt', '<u2', (4, 64))])
assert _INFO_DTYPE.itemsize == C.sizeof(_lib.JpegInfo)


def entropy_decode(images, pinned=True):
  """images: list of bytes (or a RecordBytes), all with the same geometry.  Returns (geometry info, int16 coefficients
  [B, coef_count] (pinned host tensor), uint16 quantisation tables [B, 4, 64])."""
  b = len(images)
  if b == 0:
    raise ValueError('empty batch')
  keep = []
  ptrs, lens = _pointer_arrays(images, keep)
  geom = _lib.JpegInfo()
  if _lib.lib().t2r_jpeg_parse(ptrs[0], lens[0], C.byref(geom)) != 0:
    raise UnsupportedJpeg(_lib.last_error())
  stride = int(geom.coef_count)
  coef = torch.empty((b, stride), dtype=torch.int16, pin_memory=pinned and torch.cuda.is_available())
  infos = (_lib.JpegInfo * b)()
  rc = _lib.lib().t2r_jpeg_entropy_decode_batch(ptrs, lens, b, infos, coef.data_ptr(), stride)
  del keep
  if rc != 0:
    raise UnsupportedJpeg(_lib.last_error())
  table = np.frombuffer(infos, dtype=_INFO_DTYPE)
  same = ((table['width'] == geom.width) & (table['height'] == geom.height) & (table['ncomp'] == geom.ncomp) &
          (table['h'] == table['h'][0]).all(1) & (table['v'] == table['v'][0]).all(1))
  if not same.all():
    i = int(np.argmin(same))
    inf = infos[i]
    raise UnsupportedJpeg('image %d: %dx%d, %d components, sampling %s/%s differs from the first image of the batch '
                          '(%dx%d, %d, %s/%s)' % (i, inf.width, inf.height, inf.ncomp, list(inf.h), list(inf.v),
                                                  geom.width, geom.height, geom.ncomp, list(geom.h), list(geom.v)))
  qt = table['qt'].copy()
  # table ids may differ between images: normalise to the first image's component -> table mapping
  odd = np.nonzero((table['tq'][:, :geom.ncomp] != table['tq'][0, :geom.ncomp]).any(1))[0]
  for i in odd:
    for c in range(geom.ncomp):
      if table['tq'][i, c] != geom.tq[c]:
        qt[i, geom.tq[c]] = table['qt'][i, table['tq'][i, c]]
  return geom, coef, torch.from_numpy(qt.view(np.int16))


def decode_batch(images, channels=3, device='cuda'):
  """list of JPEG byte strings (same geometry) -> uint8 CUDA tensor [B, H, W, channels]."""
  geom, coef, qt = entropy_decode(images)
  if geom.ncomp == 3 and geom.hmax == 2 and geom.width <= 4:
    # libjpeg switches from the fancy to the replicating upsampler when the chroma planes are <= 2 samples wide; the
    # device kernels implement the fancy filters only, the host decoder covers this corner
    raise UnsupportedJpeg('chroma planes of a %d pixel wide frame are upsampled by replication' % geom.width)
  dev = torch.device(device)
  if dev.type != 'cuda':
    raise _lib.T2RError('jpeg.decode_batch: the device half has no CPU path (device=%s)' % device)
  b = len(images)
  coef_d = coef.to(dev, non_blocking=True)
  qt_d = qt.to(dev, non_blocking=True)
  planes = torch.empty((b, int(geom.coef_count)), dtype=torch.uint8, device=dev)
  out = torch.empty((b, geom.height, geom.width, channels), dtype=torch.uint8, device=dev)
  stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
  _lib.call('t2r_jpeg_idct_color', C.c_void_p(coef_d.data_ptr()), C.c_void_p(qt_d.data_ptr()), C.byref(geom),
            C.c_void_p(planes.data_ptr()), C.c_void_p(out.data_ptr()), b, int(geom.coef_count), channels, stream)
  return out


def decode_batch_host(images, height, width, channels=3):
  """list of baseline JPEG byte strings, all height x width -> numpy uint8 [B, height, width, channels], decoded
  completely on host threads (t2r_jpeg_decode_host_batch): the same integer arithmetic as the device half, bit-identical
  with libjpeg-turbo.  Raises UnsupportedJpeg for streams outside the supported subset or of another size."""
  b = len(images)
  if b == 0:
    raise ValueError('empty batch')
  out = np.empty((b, height, width, channels), np.uint8)
  keep = []
  ptrs, lens = _pointer_arrays(images, keep)
  rc = _lib.lib().t2r_jpeg_decode_host_batch(ptrs, lens, b, height, width, channels, out.ctypes.data)
  del keep
  if rc != 0:
    raise UnsupportedJpeg(_lib.last_error())
  return out
