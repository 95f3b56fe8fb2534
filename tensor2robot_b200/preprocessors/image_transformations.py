"""Image transformations with the reference's function surface
(preprocessors/image_transformations.py:25-459), executed by the fused B200 kernels.

Random parameters are drawn on the host with the granularity of the reference - ONE draw per call,
shared by every image of the batch and every tensor of the list - and handed to the kernels as
parameters (TF's Philox streams cannot be reproduced, parity is on given parameters).
Crops of uint8 images stay uint8 views (no copy); conversion + photometric distortion + clip of a
cropped uint8 batch is ONE kernel (`convert_and_distort`).
"""
import numpy as np
import torch

from tensor2robot_b200.preprocessors import image_ops

_RNG = np.random.RandomState()


def seed(value):
  """Seeds the host RNG that draws crop offsets / distortion parameters."""
  global _RNG
  _RNG = np.random.RandomState(value)


def _check_shapes(images, input_shape, target_shape):
  if len(input_shape) != 3:
    raise ValueError('The input shape has to be of the form (height, width, channels) but has len {}'.format(
        len(input_shape)))
  if len(target_shape) != 2:
    raise ValueError('The target shape has to be of the form (height, width) but has len {}'.format(
        len(target_shape)))
  if input_shape[0] < target_shape[0] or input_shape[1] < target_shape[1]:
    raise ValueError('The target shape {} has to be smaller than the input shape {}'.format(
        target_shape, input_shape[:2]))
  for image in images:
    if tuple(image.shape[-3:]) != tuple(input_shape):
      raise ValueError('image shape {} does not match the declared input shape {}'.format(
          tuple(image.shape), tuple(input_shape)))


def RandomCropImages(images, input_shape, target_shape):  # pylint: disable=invalid-name
  """One random offset pair for the whole list (image_transformations.py:25-59)."""
  _check_shapes(images, input_shape, target_shape)
  max_y, max_x = input_shape[0] - target_shape[0], input_shape[1] - target_shape[1]
  oy = int(_RNG.randint(0, max_y + 1))
  ox = int(_RNG.randint(0, max_x + 1))
  return [img[..., oy:oy + target_shape[0], ox:ox + target_shape[1], :] for img in images]


def CenterCropImages(images, input_shape, target_shape):  # pylint: disable=invalid-name
  """offset = (in - out) // 2 (image_transformations.py:62-101)."""
  _check_shapes(images, input_shape, target_shape)
  oy = (input_shape[0] - target_shape[0]) // 2
  ox = (input_shape[1] - target_shape[1]) // 2
  return [img[..., oy:oy + target_shape[0], ox:ox + target_shape[1], :] for img in images]


def CustomCropImages(images, input_shape, target_shape, target_locations):  # pylint: disable=invalid-name
  """Crops around given centres, moved so the window stays inside the image (image_transformations.py:104-172).
  target_locations: one entry per image tensor, either a [batch, 2] array of (y, x) centres - one window per batch
  element, gathered into a new tensor - or a single (y, x) pair shared by the batch (a view, no copy).
  Note: the reference clamps y / x as documented but hands them to tf.image.extract_glimpse in (x, y) order; the two
  only coincide for equal coordinates (its tests use [10, 10]).  The documented (y, x) meaning is implemented here."""
  if len(input_shape) != 3:
    raise ValueError('The input shape has to be of the form (height, width, channels) but has len {}'.format(
        len(input_shape)))
  if len(target_shape) != 2:
    raise ValueError('The target shape has to be of the form (height, width) but has len {}'.format(
        len(target_shape)))
  if len(target_locations) != len(images):
    raise ValueError('There should be one target location per image. Found {} images for {} locations'.format(
        len(images), len(target_locations)))
  if input_shape[0] == target_shape[0] and input_shape[1] == target_shape[1]:
    return list(images)
  if input_shape[0] < target_shape[0] or input_shape[1] < target_shape[1]:
    raise ValueError('The target shape {} is larger than the input image size {}'.format(target_shape, input_shape[:2]))
  _check_shapes(images, input_shape, target_shape)
  th, tw = target_shape

  def window(loc):
    y = int(np.clip(loc[0], th // 2, input_shape[0] - th // 2))
    x = int(np.clip(loc[1], tw // 2, input_shape[1] - tw // 2))
    return y - th // 2, x - tw // 2

  out = []
  for img, loc in zip(images, target_locations):
    loc = np.asarray(loc.cpu() if torch.is_tensor(loc) else loc)
    if loc.ndim == 1:
      if loc.shape[0] != 2:
        raise ValueError('Target locations have to be of the form (y, x): {}'.format(loc))
      oy, ox = window(loc)
      out.append(img[..., oy:oy + th, ox:ox + tw, :])
      continue
    if loc.ndim != 2 or loc.shape[1] != 2 or loc.shape[0] != img.shape[0]:
      raise ValueError('Target locations have to be of shape [batch, 2], got {}'.format(loc.shape))
    crops = []
    for b in range(img.shape[0]):
      oy, ox = window(loc[b])
      crops.append(img[b, ..., oy:oy + th, ox:ox + tw, :])
    out.append(torch.stack(crops, 0))
  return out


def draw_photometric_params(random_brightness=False, max_delta_brightness=0.125, random_saturation=False,
                            lower_saturation=0.5, upper_saturation=1.5, random_hue=False, max_delta_hue=0.2,
                            random_contrast=False, lower_contrast=0.5, upper_contrast=1.5, random_noise_level=0.0,
                            random_noise_apply_probability=0.5):
  """The scalar draws of ApplyPhotometricImageDistortions (image_transformations.py:219-258)."""
  p = {'brightness_delta': 0.0, 'saturation_scale': 1.0, 'hue_delta': 0.0, 'contrast_scale': 1.0,
       'noise_stddev': 0.0, 'noise_seed': 0}
  if random_brightness:
    p['brightness_delta'] = float(_RNG.uniform(-max_delta_brightness, max_delta_brightness))
  if random_saturation:
    p['saturation_scale'] = float(_RNG.uniform(lower_saturation, upper_saturation))
  if random_hue:
    p['hue_delta'] = float(_RNG.uniform(-max_delta_hue, max_delta_hue))
  if random_contrast:
    p['contrast_scale'] = float(_RNG.uniform(lower_contrast, upper_contrast))
  if random_noise_level:
    # tf.cond(uniform > p, image, image + noise): noise is applied when the draw is <= p (:256-258)
    if not _RNG.uniform() > random_noise_apply_probability:
      p['noise_stddev'] = float(random_noise_level)
      p['noise_seed'] = int(_RNG.randint(0, 2**31 - 1))
  return p


def convert_and_distort(image_u8, params=None, out_dtype=torch.bfloat16):
  """uint8 [N,h,w,3] (any strided crop view) -> [0,1] floats with the given photometric parameters,
  clipped: convert_image_dtype + ApplyPhotometricImageDistortions in one kernel."""
  n = image_u8.shape[0]
  base = image_u8
  oy = ox = 0
  if not image_u8.is_contiguous() and image_u8._base is not None and image_u8._base.dim() == 4:
    # a crop view: hand the parent buffer and the offsets to the kernel instead of copying
    base = image_u8._base
    off = image_u8.storage_offset() - base.storage_offset()
    row = base.shape[2] * base.shape[3]
    oy, ox = (off % (base.shape[1] * row)) // row, (off % row) // base.shape[3]
  elif not image_u8.is_contiguous():
    base = image_u8.contiguous()
  rec = image_ops.identity_params(n, oy, ox)
  if params:
    for k in ('brightness_delta', 'saturation_scale', 'hue_delta', 'contrast_scale', 'noise_stddev'):
      rec[k] = params.get(k, rec[k][0])
  seed_value = params.get('noise_seed', 0) if params else 0
  return image_ops.crop_convert_distort(base, tuple(image_u8.shape[1:3]), rec, out_dtype, seed_value, 0)


def distort_float(image_f32, params=None):
  """The same distortions on an already converted float CUDA image [N,h,w,3] (t2r_distort_f32)."""
  n = image_f32.shape[0]
  rec = image_ops.identity_params(n)
  if params:
    for k in ('brightness_delta', 'saturation_scale', 'hue_delta', 'contrast_scale', 'noise_stddev'):
      rec[k] = params.get(k, rec[k][0])
  x = image_f32 if image_f32.dtype == torch.float32 else image_f32.float()
  y = image_ops.distort_f32(x.contiguous(), rec, params.get('noise_seed', 0) if params else 0, 0)
  return y if image_f32.dtype == torch.float32 else y.to(image_f32.dtype)


def ApplyPhotometricImageDistortions(images, **kwargs):  # pylint: disable=invalid-name
  """images: list of float CUDA tensors [B,h,w,3] in [0,1] or of uint8 crops.  One parameter draw for
  the whole list; brightness, saturation, hue, contrast, noise, then clip (:176-264)."""
  params = draw_photometric_params(**kwargs)
  out = []
  for image in images:
    if image.dtype == torch.uint8:
      out.append(convert_and_distort(image, params))
    else:
      out.append(distort_float(image, params))
  return out


def ApplyRandomFlips(images):  # pylint: disable=invalid-name
  """One left-right and one up-down coin per call, applied to the whole batch tensor (:387-399)."""
  if _RNG.uniform() < 0.5:
    images = images.flip(-2)
  if _RNG.uniform() < 0.5:
    images = images.flip(-3)
  return images


def draw_photometric_params_parallel(batch, **kwargs):
  """One independent parameter draw per image (ApplyPhotometricImageDistortionsParallel: tf.map_fn over the batch,
  image_transformations.py:316-362).  Returns (DISTORT_DTYPE records [batch], noise seed)."""
  rec = image_ops.identity_params(batch)
  seed_value = 0
  for i in range(batch):
    p = draw_photometric_params(**kwargs)
    for k in ('brightness_delta', 'saturation_scale', 'hue_delta', 'contrast_scale', 'noise_stddev'):
      rec[k][i] = p[k]
    seed_value = seed_value or p['noise_seed']      # one Philox key per call; the counter separates the images
  return rec, seed_value


def ApplyPhotometricImageDistortionsParallel(images, custom_distortion_fn=None, **kwargs):  # pylint: disable=invalid-name
  """images: ONE CUDA tensor [B,h,w,3] (uint8 or float in [0,1]); every image gets its own random brightness /
  saturation / hue / contrast / noise draw, then the clip - still a single kernel launch, the per-image values are
  rows of the parameter table (:268-362)."""
  if custom_distortion_fn is not None:
    raise NotImplementedError('custom_distortion_fn would run arbitrary per-image code between kernel stages')
  rec, seed_value = draw_photometric_params_parallel(images.shape[0], **kwargs)
  if images.dtype == torch.uint8:
    return image_ops.crop_convert_distort(images.contiguous(), tuple(images.shape[1:3]), rec, torch.float32, seed_value, 0)
  x = images if images.dtype == torch.float32 else images.float()
  y = image_ops.distort_f32(x.contiguous(), rec, seed_value, 0)
  return y if images.dtype == torch.float32 else y.to(images.dtype)


def ApplyPhotometricImageDistortionsCheap(images):  # pylint: disable=invalid-name
  """Per-channel random gamma correction, gamma ~ U[0.5, 1.5) drawn once per channel for the whole batch
  (:365-384).  images: float CUDA tensor [B,h,w,3] in (0, 1)."""
  gammas = [float(_RNG.uniform(0.5, 1.5)) for _ in range(images.shape[-1])]
  x = images if images.dtype == torch.float32 else images.float()
  y = image_ops.channel_gamma(x, gammas)
  return y if images.dtype == torch.float32 else y.to(images.dtype)


def ApplyDepthImageDistortions(depth_images, random_noise_level=0.05, random_noise_apply_probability=0.5,  # pylint: disable=invalid-name
                               scaling_noise=True, gamma_shape=1000.0, gamma_scale_inverse=1000.0,
                               min_depth_allowed=0.25, max_depth_allowed=2.5):
  """depth_images: list of float CUDA tensors [B,h,w,1].  Per list entry: with probability
  `random_noise_apply_probability` the tensor becomes alpha * depth + N(0, random_noise_level), alpha ~
  Gamma(gamma_shape, rate gamma_scale_inverse); every tensor is clipped to the allowed depth range (:403-459).
  (The reference leaves alpha undefined when scaling_noise is False; alpha = 1 here.)"""
  assert depth_images[0].shape[-1] == 1
  out = []
  for image in depth_images:
    alpha, sigma, seed_value = 1.0, 0.0, 0
    if random_noise_level:
      seed_value = int(_RNG.randint(0, 2**31 - 1))
      drawn_alpha = float(_RNG.gamma(gamma_shape, 1.0 / gamma_scale_inverse)) if scaling_noise else 1.0
      # tf.cond(uniform > p, image, alpha * image + noise): the distortion is applied when the draw is <= p
      if not _RNG.uniform() > random_noise_apply_probability:
        alpha, sigma = drawn_alpha, float(random_noise_level)
    x = image if image.dtype == torch.float32 else image.float()
    y = image_ops.depth_distort(x, alpha, sigma, min_depth_allowed, max_depth_allowed, seed_value, 0)
    out.append(y if image.dtype == torch.float32 else y.to(image.dtype))
  return out
