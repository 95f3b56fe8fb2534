"""Device-side image preprocessing primitives (fused crop + convert + photometric distortion, legacy
bilinear resize) over the C-ABI.  The preprocessor classes in image_transformations.py /
distortion.py draw the random parameters on the host exactly where the reference draws them and
hand them to these kernels (SURVEY 7 'hard parts': RNG streams cannot match TF's, so parity is on
*given* parameters).
"""
import ctypes as C

import numpy as np
import torch

from tensor2robot_b200 import _lib

# numpy mirror of T2RDistortParams (include/t2r_b200.h)
DISTORT_DTYPE = np.dtype([('brightness_delta', np.float32), ('saturation_scale', np.float32),
                          ('hue_delta', np.float32), ('contrast_scale', np.float32),
                          ('noise_stddev', np.float32), ('crop_y', np.int32), ('crop_x', np.int32),
                          ('reserved', np.int32)])
assert DISTORT_DTYPE.itemsize == C.sizeof(_lib.DistortParams)


def _p(t):
  return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
  return _lib.current_stream_ptr()


def identity_params(n, crop_y=0, crop_x=0):
  """Per-image parameter records that only crop and convert."""
  p = np.zeros(n, DISTORT_DTYPE)
  p['saturation_scale'] = 1.0
  p['contrast_scale'] = 1.0
  p['crop_y'] = crop_y
  p['crop_x'] = crop_x
  return p


def crop_convert_distort(images_u8, out_hw, params, out_dtype=torch.bfloat16, seed=0, offset=0):
  """images_u8: CUDA uint8 [N,H,W,3]; params: numpy DISTORT_DTYPE [N].  Returns [N,h,w,3] in [0,1]."""
  if not images_u8.is_cuda or images_u8.dtype != torch.uint8:
    raise _lib.T2RError('crop_convert_distort needs a CUDA uint8 tensor (no CPU path)')
  n, hh, ww, c = images_u8.shape
  if c != 3:
    raise ValueError('photometric distortions are defined for 3-channel images, got %d channels' % c)
  h, w = out_hw
  params = np.ascontiguousarray(params, dtype=DISTORT_DTYPE)
  if params.shape != (n,):
    raise ValueError('need one parameter record per image')
  if (params['crop_y'] < 0).any() or (params['crop_y'] + h > hh).any() or (params['crop_x'] < 0).any() or (
      params['crop_x'] + w > ww).any():
    raise ValueError('crop window outside the %dx%d image' % (hh, ww))
  dev_params = torch.from_numpy(params.view(np.uint8).reshape(n, -1)).to(images_u8.device, non_blocking=True)
  use_contrast = bool((params['contrast_scale'] != 1.0).any())
  out = torch.empty((n, h, w, 3), dtype=out_dtype, device=images_u8.device)
  chan_mean = torch.empty((n, 3), dtype=torch.float32, device=images_u8.device) if use_contrast else None
  _lib.call('t2r_crop_convert_distort', _p(images_u8.contiguous()), _p(out), _p(dev_params), _p(chan_mean), n, hh,
            ww, h, w, 1 if out_dtype == torch.float32 else 0, 1 if use_contrast else 0, int(seed), int(offset),
            _stream())
  return out


def distort_f32(images_f32, params, seed=0, offset=0):
  """Photometric distortions on an already converted CUDA float32 image batch [N,H,W,3] in [0,1]
  (the BC-Z order: distort after the resize, preprocessors/distortion.py:56-107)."""
  if not images_f32.is_cuda or images_f32.dtype != torch.float32:
    raise _lib.T2RError('distort_f32 needs a CUDA float32 tensor (no CPU path)')
  n, h, w, c = images_f32.shape
  if c != 3:
    raise ValueError('photometric distortions are defined for 3-channel images, got %d channels' % c)
  params = np.ascontiguousarray(params, dtype=DISTORT_DTYPE)
  if params.shape != (n,) or params['crop_y'].any() or params['crop_x'].any():
    raise ValueError('need one parameter record per image, without a crop')
  dev_params = torch.from_numpy(params.view(np.uint8).reshape(n, -1)).to(images_f32.device, non_blocking=True)
  use_contrast = bool((params['contrast_scale'] != 1.0).any())
  out = torch.empty_like(images_f32)
  chan_mean = torch.empty((n, 3), dtype=torch.float32, device=images_f32.device) if use_contrast else None
  _lib.call('t2r_distort_f32', _p(images_f32.contiguous()), _p(out), _p(dev_params), _p(chan_mean), n, h, w,
            1 if use_contrast else 0, int(seed), int(offset), _stream())
  return out


def resize_bilinear_legacy(images_f32, out_hw):
  """tf.image.resize_images(BILINEAR) with TF1 legacy sampling; CUDA float32 NHWC."""
  if not images_f32.is_cuda or images_f32.dtype != torch.float32:
    raise _lib.T2RError('resize_bilinear_legacy needs a CUDA float32 tensor (no CPU path)')
  n, hh, ww, c = images_f32.shape
  h, w = out_hw
  out = torch.empty((n, h, w, c), dtype=torch.float32, device=images_f32.device)
  _lib.call('t2r_resize_bilinear_legacy', _p(images_f32.contiguous()), _p(out), n, hh, ww, c, h, w, _stream())
  return out


def channel_gamma(images_f32, gammas):
  """dst = src ** gamma_c per channel (t2r_channel_gamma_f32): fp32 CUDA [..., C], C <= 4."""
  if not images_f32.is_cuda or images_f32.dtype != torch.float32:
    raise _lib.T2RError('channel_gamma needs a CUDA float32 tensor (no CPU path)')
  c = images_f32.shape[-1]
  if len(gammas) != c or c > 4:
    raise ValueError('need one gamma per channel (at most 4 channels)')
  g = [float(v) for v in gammas] + [1.0] * (4 - c)
  x = images_f32.contiguous()
  out = torch.empty_like(x)
  _lib.call('t2r_channel_gamma_f32', _p(x), _p(out), x.numel(), c, g[0], g[1], g[2], g[3], _stream())
  return out


def depth_distort(depth_f32, alpha, noise_stddev, min_depth, max_depth, seed=0, offset=0):
  """clip(alpha * x + N(0, noise_stddev), min_depth, max_depth) on a CUDA float32 tensor (t2r_depth_distort_f32)."""
  if not depth_f32.is_cuda or depth_f32.dtype != torch.float32:
    raise _lib.T2RError('depth_distort needs a CUDA float32 tensor (no CPU path)')
  x = depth_f32.contiguous()
  out = torch.empty_like(x)
  _lib.call('t2r_depth_distort_f32', _p(x), _p(out), x.numel(), float(alpha), float(noise_stddev), float(min_depth),
            float(max_depth), int(seed), int(offset), _stream())
  return out
