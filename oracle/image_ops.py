"""ORACLE (test infrastructure, never imported by the product): numpy restatement of the image ops
the reference's preprocessors call.

  crop / convert / distort / clip      preprocessors/image_transformations.py:25-101,176-264;
                                       research/qtopt/t2r_models.py:297-308
  resize_bilinear_legacy               preprocessors/distortion.py:56-107 (tf.image.resize_images)

The arithmetic lives in TensorFlow 1.x kernels (absent from /root/reference); restated from TF 1.15:
  convert_image_dtype(uint8->float32)  x * (1/255)            (image_ops_impl.py: multiply by scale)
  adjust_brightness                    x + delta
  adjust_saturation / adjust_hue       core/kernels/adjust_saturation_op.cc / adjust_hue_op.cc via
                                       rgb_to_hsv / hsv_to_rgb (core/kernels/image/adjust_hsv_gpu.cu.h
                                       formulas, restated below)
  adjust_contrast                      (x - mean_c) * f + mean_c, mean over H,W per image & channel
  resize_images(BILINEAR)              align_corners=False, legacy sampling src = dst * in/out
PARITY UNPINNED: the reference's tests only check that distortion "changed something"
(image_transformations_test.py:162-199).  All arithmetic in float32 to mirror the kernels.
"""
import numpy as np

F = np.float32


def crop(images, oy, ox, h, w):
  return images[:, oy:oy + h, ox:ox + w, :]


def convert_image_dtype_f32(images_u8):
  return images_u8.astype(F) * F(1.0 / 255.0)


def rgb_to_hsv(rgb):
  r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
  mx = np.maximum(r, np.maximum(g, b))
  mn = np.minimum(r, np.minimum(g, b))
  rng = mx - mn
  v = mx
  with np.errstate(divide='ignore', invalid='ignore'):
    s = np.where(mx > 0, rng / mx, F(0)).astype(F)
    norm = (F(1.0) / (F(6.0) * rng)).astype(F)
    h = np.where(r == mx, norm * (g - b),
                 np.where(g == mx, norm * (b - r) + F(2.0 / 6.0), norm * (r - g) + F(4.0 / 6.0))).astype(F)
  h = np.where(rng > 0, h, F(0)).astype(F)
  h = np.where(h < 0, h + F(1.0), h).astype(F)
  return h, s, v


def hsv_to_rgb(h, s, v):
  c = (s * v).astype(F)
  m = (v - c).astype(F)
  dh = (h * F(6.0)).astype(F)
  cat = dh.astype(np.int32)
  fm = dh.copy()
  fm = np.where(fm <= 0, fm + F(2.0), fm).astype(F)
  while (fm >= 2).any():
    fm = np.where(fm >= 2, fm - F(2.0), fm).astype(F)
  x = (c * (F(1.0) - np.abs(fm - F(1.0)))).astype(F)
  z = np.zeros_like(c)
  table = {0: (c, x, z), 1: (x, c, z), 2: (z, c, x), 3: (z, x, c), 4: (x, z, c), 5: (c, z, x)}
  out = np.zeros(h.shape + (3,), F)
  for k, (rr, gg, bb) in table.items():
    sel = cat == k
    out[..., 0] = np.where(sel, rr, out[..., 0])
    out[..., 1] = np.where(sel, gg, out[..., 1])
    out[..., 2] = np.where(sel, bb, out[..., 2])
  return (out + m[..., None]).astype(F)


def distort(images_f32, brightness_delta=0.0, saturation_scale=1.0, hue_delta=0.0, contrast_scale=1.0, noise=None):
  """images [N,h,w,3] float32; scalar parameters shared by the batch (the reference draws one per call);
  noise: optional array added before the clip."""
  x = images_f32.astype(F)
  if brightness_delta != 0.0:
    x = (x + F(brightness_delta)).astype(F)
  if saturation_scale != 1.0:
    h, s, v = rgb_to_hsv(x)
    s = np.clip(s * F(saturation_scale), 0, 1).astype(F)
    x = hsv_to_rgb(h, s, v)
  if hue_delta != 0.0:
    h, s, v = rgb_to_hsv(x)
    h = (h + F(hue_delta)).astype(F)
    h = (h - np.floor(h)).astype(F)
    x = hsv_to_rgb(h, s, v)
  if contrast_scale != 1.0:
    mean = x.astype(np.float64).mean((1, 2), keepdims=True).astype(F)
    x = ((x - mean) * F(contrast_scale) + mean).astype(F)
  if noise is not None:
    x = (x + noise.astype(F)).astype(F)
  return np.clip(x, 0.0, 1.0).astype(F)


def resize_bilinear_legacy(images, out_h, out_w):
  n, h, w, c = images.shape
  sy, sx = F(h) / F(out_h), F(w) / F(out_w)
  ys = (np.arange(out_h, dtype=F) * sy).astype(F)
  xs = (np.arange(out_w, dtype=F) * sx).astype(F)
  y0 = np.floor(ys).astype(np.int64)
  x0 = np.floor(xs).astype(np.int64)
  y1 = np.minimum(y0 + 1, h - 1)
  x1 = np.minimum(x0 + 1, w - 1)
  ly = (ys - y0.astype(F)).astype(F)[None, :, None, None]
  lx = (xs - x0.astype(F)).astype(F)[None, None, :, None]
  img = images.astype(F)
  tl, tr = img[:, y0][:, :, x0], img[:, y0][:, :, x1]
  bl, br = img[:, y1][:, :, x0], img[:, y1][:, :, x1]
  top = tl + (tr - tl) * lx
  bot = bl + (br - bl) * lx
  return (top + (bot - top) * ly).astype(F)
