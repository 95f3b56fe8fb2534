"""Baseline JPEG decode parity (CPU part): the numpy oracle (oracle/jpeg.py, a restatement of libjpeg's
ISLOW IDCT + fancy upsampling + YCbCr tables, which is what tf.image.decode_image runs for
utils/tfdata.py:426-484) is pinned bit-exactly against libjpeg-turbo through PIL, on the reference's own
fixture images and on synthetic JPEGs covering 4:4:4 / 4:2:2 / 4:2:0, greyscale, odd sizes, restart intervals."""
import io
import os

import numpy as np
import pytest
from PIL import Image

from oracle import jpeg as oracle_jpeg
from oracle import tfrecord as oracle_tfrecord

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, 'golden', 'pose_env_test_data.tfrecord')


def _pil(data, mode):
  return np.array(Image.open(io.BytesIO(data)).convert(mode) if mode == 'RGB' else Image.open(io.BytesIO(data)))


def _picture(h, w, seed):
  rng = np.random.RandomState(seed)
  yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
  img = np.stack([127 + 120 * np.sin(xx / 7.0 + seed), 127 + 120 * np.cos(yy / 5.0), (xx * 3 + yy * 2) % 256], -1)
  img += rng.uniform(-25, 25, img.shape)
  img[h // 3:h // 2, w // 4:w // 2] = rng.randint(0, 256, 3)    # a hard-edged patch
  return np.clip(img, 0, 255).astype(np.uint8)


def _encode(img, **kw):
  buf = io.BytesIO()
  Image.fromarray(img).save(buf, format='JPEG', **kw)
  return buf.getvalue()


def test_oracle_matches_libjpeg_on_reference_fixture():
  records = oracle_tfrecord.read_tfrecords(FIXTURE)
  n = 0
  for rec in records[:4]:
    ex = oracle_tfrecord.parse_example(rec)
    for key, (kind, values) in ex.items():
      if kind == 'bytes' and values and values[0][:2] == b'\xff\xd8':
        want = _pil(values[0], 'RGB')
        got = oracle_jpeg.decode(values[0])
        assert got.shape == want.shape
        np.testing.assert_array_equal(got, want)
        n += 1
  assert n >= 4


@pytest.mark.parametrize('h,w,subsampling,quality', [
    (64, 64, 0, 90), (48, 80, 1, 75), (64, 96, 2, 85), (37, 29, 2, 60), (33, 50, 1, 95), (17, 23, 0, 50),
    (100, 75, 2, 30)])
def test_oracle_matches_libjpeg_colour(h, w, subsampling, quality):
  data = _encode(_picture(h, w, h + w), quality=quality, subsampling=subsampling)
  np.testing.assert_array_equal(oracle_jpeg.decode(data), _pil(data, 'RGB'))


def test_oracle_matches_libjpeg_grey_and_restart_markers():
  img = _picture(40, 56, 3)
  grey = _encode(img[..., 0], quality=80)
  want = np.array(Image.open(io.BytesIO(grey)))
  np.testing.assert_array_equal(oracle_jpeg.decode(grey, channels=1)[..., 0], want)
  np.testing.assert_array_equal(oracle_jpeg.decode(grey, channels=3), np.repeat(want[..., None], 3, -1))
  try:
    data = _encode(img, quality=70, subsampling=2, restart_marker_blocks=3)
  except TypeError:
    pytest.skip('this Pillow cannot write restart markers')
  assert oracle_jpeg.parse_headers(data)['restart_interval'] == 3
  np.testing.assert_array_equal(oracle_jpeg.decode(data), _pil(data, 'RGB'))


def test_unsupported_and_corrupt_inputs_raise():
  with pytest.raises(oracle_jpeg.JpegError):
    oracle_jpeg.decode(b'not a jpeg at all')
  prog = _encode(_picture(32, 32, 1), progressive=True)
  with pytest.raises(oracle_jpeg.JpegError):
    oracle_jpeg.decode(prog)
