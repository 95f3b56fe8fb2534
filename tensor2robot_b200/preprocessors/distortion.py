"""BC-Z style image preprocessing: convert, crop, legacy bilinear resize, distort
(preprocessors/distortion.py:56-141)."""
import torch

from tensor2robot_b200.preprocessors import image_ops
from tensor2robot_b200.preprocessors import image_transformations


def crop_image(img, mode, input_size=(512, 640), target_size=(472, 472)):
  """Random crop in TRAIN, centre crop otherwise (distortion.py:110-141)."""
  if input_size == target_size:
    return img
  input_shape = tuple(input_size) + (img.shape[-1],)
  if mode == 'train':
    return image_transformations.RandomCropImages([img], input_shape, target_size)[0]
  return image_transformations.CenterCropImages([img], input_shape, target_size)[0]


def preprocess_image(image, mode, is_sequence, input_size, target_size, crop_size=None,
                     image_distortion_fn=None):
  """uint8 [B,(T,)H,W,3] -> float32 [B,(T,)h,w,3]: to-float, crop, TF1-legacy bilinear resize,
  distortion (distortion.py:56-107).  image_distortion_fn: kwargs dict for
  ApplyPhotometricImageDistortions (TRAIN only) or None."""
  shape = image.shape
  if is_sequence:
    image = image.reshape((-1,) + tuple(shape[2:]))
  crop = crop_size or input_size
  cropped = crop_image(image, mode, input_size, crop)
  params = None
  if image_distortion_fn and mode == 'train':
    params = image_transformations.draw_photometric_params(**image_distortion_fn)
  if tuple(crop) == tuple(target_size):
    out = image_transformations.convert_and_distort(cropped, params, torch.float32)
  else:
    # the reference distorts AFTER the resize (distortion.py:95-104): convert -> resize -> distort
    f = image_transformations.convert_and_distort(cropped, None, torch.float32)
    out = image_ops.resize_bilinear_legacy(f, tuple(target_size))
    out = image_transformations.distort_float(out, params)   # also the final clip to [0, 1]
  if is_sequence:
    out = out.reshape(tuple(shape[:2]) + tuple(out.shape[1:]))
  return out
