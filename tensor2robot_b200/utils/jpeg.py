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


def parse(data):
  """Header fields of one JPEG stream as a `_lib.JpegInfo`."""
  info = _lib.JpegInfo()
  buf = (C.c_char * len(data)).from_buffer_copy(data)
  rc = _lib.lib().t2r_jpeg_parse(C.addressof(buf), len(data), C.byref(info))
  if rc != 0:
    raise UnsupportedJpeg(_lib.last_error())
  return info


def entropy_decode(images, pinned=True):
  """images: list of bytes, all with the same geometry.  Returns (geometry info, int16 coefficients
  [B, coef_count] (pinned host tensor), uint16 quantisation tables [B, 4, 64])."""
  b = len(images)
  if b == 0:
    raise ValueError('empty batch')
  geom = parse(images[0])
  stride = int(geom.coef_count)
  coef = torch.empty((b, stride), dtype=torch.int16, pin_memory=pinned and torch.cuda.is_available())
  infos = (_lib.JpegInfo * b)()
  ptrs = (C.c_void_p * b)()
  lens = (C.c_uint64 * b)()
  keep = []
  for i, img in enumerate(images):
    buf = (C.c_char * len(img)).from_buffer_copy(img)
    keep.append(buf)
    ptrs[i], lens[i] = C.addressof(buf), len(img)
  rc = _lib.lib().t2r_jpeg_entropy_decode_batch(ptrs, lens, b, infos, coef.data_ptr(), stride)
  if rc != 0:
    raise UnsupportedJpeg(_lib.last_error())
  qt = np.zeros((b, 4, 64), np.uint16)
  for i in range(b):
    inf = infos[i]
    same = (inf.width == geom.width and inf.height == geom.height and inf.ncomp == geom.ncomp and
            list(inf.h) == list(geom.h) and list(inf.v) == list(geom.v))
    if not same:
      raise UnsupportedJpeg('image %d: %dx%d, %d components, sampling %s/%s differs from the first image of the batch '
                            '(%dx%d, %d, %s/%s)' % (i, inf.width, inf.height, inf.ncomp, list(inf.h), list(inf.v),
                                                    geom.width, geom.height, geom.ncomp, list(geom.h), list(geom.v)))
    qt[i] = np.ctypeslib.as_array(inf.qt)
    # table ids may differ between images: normalise to the first image's component -> table mapping
    for c in range(inf.ncomp):
      if inf.tq[c] != geom.tq[c]:
        qt[i, geom.tq[c]] = np.ctypeslib.as_array(inf.qt)[inf.tq[c]]
  del keep
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
  ptrs = (C.c_void_p * b)()
  lens = (C.c_uint64 * b)()
  keep = []
  for i, img in enumerate(images):
    buf = (C.c_char * len(img)).from_buffer_copy(img)
    keep.append(buf)
    ptrs[i], lens[i] = C.addressof(buf), len(img)
  rc = _lib.lib().t2r_jpeg_decode_host_batch(ptrs, lens, b, height, width, channels, out.ctypes.data)
  del keep
  if rc != 0:
    raise UnsupportedJpeg(_lib.last_error())
  return out
