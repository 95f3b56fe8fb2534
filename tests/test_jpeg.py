"""Baseline JPEG decode parity (CPU part): the numpy oracle (oracle/jpeg.py, a restatement of libjpeg's
ISLOW IDCT + fancy upsampling + YCbCr tables, which is what tf.image.decode_image runs for
utils/tfdata.py:426-484) is pinned bit-exactly against libjpeg-turbo through PIL, on the reference's own
fixture images and on synthetic JPEGs covering 4:4:4 / 4:2:2 / 4:2:0, greyscale, odd sizes, restart intervals."""
import io
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st
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


# ---- the engine's split decoder ---------------------------------------------------------------------
def _cases():
  yield 'fixture', None
  for h, w, sub, q in [(64, 64, 0, 90), (48, 80, 1, 75), (64, 96, 2, 85), (37, 29, 2, 60), (33, 50, 1, 95)]:
    yield '%dx%d_s%d' % (h, w, sub), _encode(_picture(h, w, h + w), quality=q, subsampling=sub)
  yield 'grey', _encode(_picture(40, 56, 3)[..., 0], quality=80)


def _fixture_jpegs():
  out = []
  for rec in oracle_tfrecord.read_tfrecords(FIXTURE)[:3]:
    for key, (kind, values) in oracle_tfrecord.parse_example(rec).items():
      if kind == 'bytes' and values and values[0][:2] == b'\xff\xd8':
        out.append(values[0])
  return out


def test_host_entropy_decoder_matches_oracle_coefficients():
  """csrc/jpeg_host.cc (headers + Huffman, threaded over the batch) against the oracle's T.81 Annex F walk:
  identical quantised coefficients, quantisation tables and geometry."""
  from tensor2robot_b200.utils import jpeg
  for name, data in _cases():
    batch = _fixture_jpegs() if data is None else [data, data]
    geom, coef, qt = jpeg.entropy_decode(batch, pinned=False)
    for i, img in enumerate(batch):
      info, coefs = oracle_jpeg.decode_coefficients(img)
      assert (geom.width, geom.height, geom.ncomp) == (info['width'], info['height'], len(info['comps']))
      for c, (comp, want) in enumerate(zip(info['comps'], coefs)):
        n = want.size
        got = coef[i, geom.coef_offset[c]:geom.coef_offset[c] + n].numpy().reshape(want.shape)
        np.testing.assert_array_equal(got, want, err_msg='%s image %d component %d' % (name, i, c))
        np.testing.assert_array_equal(qt[i].numpy().view(np.uint16)[geom.tq[c]], info['qt'][comp[3]])
  with pytest.raises(jpeg.UnsupportedJpeg):
    jpeg.entropy_decode([_encode(_picture(32, 32, 1), progressive=True)], pinned=False)
  with pytest.raises(jpeg.UnsupportedJpeg):
    jpeg.entropy_decode([_encode(_picture(32, 32, 1)), _encode(_picture(32, 40, 1))], pinned=False)


@pytest.mark.gpu
def test_split_decoder_matches_libjpeg_bit_exact():
  """Host Huffman + device IDCT / upsampling / colour == libjpeg-turbo (PIL) == the oracle, bit for bit."""
  import torch
  from tensor2robot_b200.utils import jpeg
  for name, data in _cases():
    batch = _fixture_jpegs() if data is None else [data, data, data]
    got = jpeg.decode_batch(batch, channels=3).cpu().numpy()
    for i, img in enumerate(batch):
      want = _pil(img, 'RGB')
      np.testing.assert_array_equal(got[i], want, err_msg='%s image %d' % (name, i))
      np.testing.assert_array_equal(got[i], oracle_jpeg.decode(img))
  # luma-only output (tf.image.decode_image(channels=1) on a colour JPEG)
  data = _encode(_picture(64, 96, 5), quality=85, subsampling=2)
  np.testing.assert_array_equal(jpeg.decode_batch([data], channels=1).cpu().numpy()[0],
                                oracle_jpeg.decode(data, channels=1))
  # replay-frame size: 512 x 640, 4:2:0
  big = _encode(_picture(512, 640, 9), quality=90, subsampling=2)
  np.testing.assert_array_equal(jpeg.decode_batch([big] * 4).cpu().numpy()[3], _pil(big, 'RGB'))


@pytest.mark.gpu
def test_parser_with_device_decoder_matches_host_decoder():
  """create_parse_tf_example_fn on the reference fixture: T2R_IMAGE_DECODER=device returns the same pixels
  (as a uint8 CUDA tensor) as the host (PIL) path."""
  from tensor2robot_b200.utils import dtypes, tensorspec_utils as utils, tfdata
  spec = utils.TensorSpecStruct(
      state=utils.TensorSpecStruct(image=utils.ExtendedTensorSpec((64, 64, 3), dtypes.uint8, 'state/image',
                                                                  data_format='jpeg')))
  records = oracle_tfrecord.read_tfrecords(FIXTURE)[:6]
  parse = tfdata.create_parse_tf_example_fn(spec)
  tfdata.set_image_decoder('host')
  host = parse(records).state.image
  tfdata.set_image_decoder('device')
  try:
    dev = parse(records).state.image
  finally:
    tfdata.set_image_decoder('auto')
  assert hasattr(dev, 'is_cuda') and dev.is_cuda and tuple(dev.shape) == host.shape
  np.testing.assert_array_equal(dev.cpu().numpy(), host)


def test_host_entropy_decoder_property():
  """Property test (hypothesis): for random small images, qualities and chroma subsamplings the C++ Huffman
  decoder returns exactly the oracle's coefficients, and the oracle's pixels equal libjpeg-turbo's."""
  import hypothesis
  from hypothesis import strategies as st
  from tensor2robot_b200.utils import jpeg

  @hypothesis.settings(max_examples=25, deadline=None)
  @hypothesis.given(st.integers(8, 40), st.integers(8, 40), st.integers(5, 100), st.sampled_from([0, 1, 2]),
                    st.integers(0, 2**31 - 1))
  def check(h, w, quality, subsampling, seed):
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    img[: h // 2] = (img[: h // 2].astype(np.int32) // 4 + 96).astype(np.uint8)      # a smoother half
    data = _encode(img, quality=quality, subsampling=subsampling)
    geom, coef, _ = jpeg.entropy_decode([data], pinned=False)
    info, coefs = oracle_jpeg.decode_coefficients(data)
    for c, want in enumerate(coefs):
      got = coef[0, geom.coef_offset[c]:geom.coef_offset[c] + want.size].numpy().reshape(want.shape)
      np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(oracle_jpeg.decode(data), _pil(data, 'RGB'))

  check()


# ---- the complete host decoder (t2r_jpeg_decode_host_batch): the 'host' image decoder of the record parser ----------
def test_host_decoder_matches_libjpeg_bit_exact():
  """Huffman + ISLOW IDCT + fancy upsampling + colour on C++ host threads == libjpeg-turbo (PIL), bit for bit, for every
  sampling layout, odd sizes, restart intervals, grey streams, luma-only output, and the whole reference fixture."""
  from tensor2robot_b200.utils import jpeg
  fixture = []
  for rec in oracle_tfrecord.read_tfrecords(FIXTURE):
    fixture.append(oracle_tfrecord.parse_example(rec)['state/image'][1][0])
  got = jpeg.decode_batch_host(fixture, 64, 64, 3)
  want = np.stack([np.asarray(Image.open(io.BytesIO(b)).convert('RGB')) for b in fixture])
  np.testing.assert_array_equal(got, want)
  for name, data in _cases():
    if data is None:
      continue
    ref = np.asarray(Image.open(io.BytesIO(data)))
    h, w = ref.shape[:2]
    rgb = ref if ref.ndim == 3 else np.repeat(ref[..., None], 3, -1)
    np.testing.assert_array_equal(jpeg.decode_batch_host([data] * 9, h, w, 3), np.stack([rgb] * 9), err_msg=name)
    if ref.ndim == 2:
      np.testing.assert_array_equal(jpeg.decode_batch_host([data], h, w, 1)[0, ..., 0], ref)
  restart = _encode(_picture(50, 70, 9), quality=70, subsampling=2, restart_marker_blocks=3)
  np.testing.assert_array_equal(jpeg.decode_batch_host([restart], 50, 70)[0], np.asarray(Image.open(io.BytesIO(restart))))


@settings(max_examples=25, deadline=None)
@given(st.integers(8, 70), st.integers(8, 70), st.sampled_from([0, 1, 2]), st.integers(20, 100), st.integers(0, 2**31 - 1))
def test_host_decoder_property(h, w, subsampling, quality, seed):
  from tensor2robot_b200.utils import jpeg
  rng = np.random.RandomState(seed)
  img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
  data = _encode(img, quality=quality, subsampling=subsampling)
  np.testing.assert_array_equal(jpeg.decode_batch_host([data], h, w)[0], np.asarray(Image.open(io.BytesIO(data))))


def test_host_decoder_rejections_and_parser_fallback():
  from tensor2robot_b200.utils import dtypes
  from tensor2robot_b200.utils import jpeg
  from tensor2robot_b200.utils import tensorspec_utils as utils
  from tensor2robot_b200.utils import tfdata
  data = _encode(_picture(48, 64, 1), quality=80, subsampling=2)
  with pytest.raises(jpeg.UnsupportedJpeg, match='expects'):
    jpeg.decode_batch_host([data], 64, 48)                       # another size
  with pytest.raises(jpeg.UnsupportedJpeg, match='premature end'):
    jpeg.decode_batch_host([data[:len(data) // 2]], 48, 64)      # truncated: tf.image.decode_image refuses it too
  with pytest.raises(jpeg.UnsupportedJpeg, match='premature end'):
    jpeg.decode_batch_host([data[:-2]], 48, 64)                  # EOI missing
  buf = io.BytesIO()
  Image.fromarray(_picture(48, 64, 1)).save(buf, format='JPEG', quality=80, progressive=True)
  progressive = buf.getvalue()
  with pytest.raises(jpeg.UnsupportedJpeg, match='baseline'):
    jpeg.decode_batch_host([progressive], 48, 64)
  # the record parser sends what the C++ decoder refuses through PIL: same pixels as PIL's own decode
  spec = utils.TensorSpecStruct(image=utils.ExtendedTensorSpec((48, 64, 3), dtypes.uint8, 'image', data_format='jpeg'))
  records = [oracle_tfrecord.make_example({'image': progressive}), oracle_tfrecord.make_example({'image': progressive})]
  parsed = tfdata.create_parse_tf_example_fn(spec)(records)
  np.testing.assert_array_equal(np.asarray(parsed.image)[1], np.asarray(Image.open(io.BytesIO(progressive))))
  mixed = [oracle_tfrecord.make_example({'image': data}), oracle_tfrecord.make_example({'image': progressive})]
  parsed = tfdata.create_parse_tf_example_fn(spec)(mixed)
  np.testing.assert_array_equal(np.asarray(parsed.image)[0], np.asarray(Image.open(io.BytesIO(data))))
  bad = [oracle_tfrecord.make_example({'image': data[:300]})]
  with pytest.raises(ValueError):
    tfdata.create_parse_tf_example_fn(spec)(bad)


def test_host_decoder_every_small_size():
  """All frame sizes 1..12 (+ a few around the MCU sizes) x 4:4:4 / 4:2:2 / 4:2:0: edge columns, odd sizes and libjpeg's
  switch to the replicating upsampler for chroma planes <= 2 samples wide (jdsample.c jinit_upsampler) - host decoder
  and oracle against libjpeg-turbo; the device decoder hands the narrow-plane corner to the host decoder."""
  from tensor2robot_b200.utils import jpeg
  rng = np.random.RandomState(0)
  sizes = list(range(1, 13)) + [15, 16, 17, 31, 32, 33]
  for h in sizes:
    for w in sizes:
      img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
      for subsampling in (0, 1, 2):
        data = _encode(img, quality=75, subsampling=subsampling)
        want = np.asarray(Image.open(io.BytesIO(data)).convert('RGB'))
        np.testing.assert_array_equal(jpeg.decode_batch_host([data], h, w)[0], want, err_msg=str((h, w, subsampling)))
        if w <= 5 and h in (1, 2, 9):
          np.testing.assert_array_equal(oracle_jpeg.decode(data), want, err_msg='oracle ' + str((h, w, subsampling)))
  narrow = _encode(rng.randint(0, 256, (16, 4, 3)).astype(np.uint8), quality=75, subsampling=2)
  with pytest.raises(jpeg.UnsupportedJpeg, match='replication'):
    jpeg.decode_batch([narrow])


def test_luma_only_output_is_the_y_plane():
  """tf.image.decode_image(channels=1) asks libjpeg for JCS_GRAYSCALE, i.e. the Y plane - not an RGB -> L conversion.
  PIL reaches the same libjpeg mode through draft('L')."""
  from tensor2robot_b200.utils import jpeg
  data = _encode(_picture(40, 56, 5), quality=85, subsampling=2)
  got = jpeg.decode_batch_host([data], 40, 56, channels=1)[0, ..., 0]
  im = Image.open(io.BytesIO(data))
  im.draft('L', im.size)
  np.testing.assert_array_equal(got, np.asarray(im))
  np.testing.assert_array_equal(got, oracle_jpeg.decode(data, channels=1)[..., 0])
  converted = np.asarray(Image.open(io.BytesIO(data)).convert('L'))
  diff = np.abs(got.astype(np.int32) - converted.astype(np.int32))
  assert 0 < diff.max() <= 8                          # close to, but not, PIL's RGB -> L formula (clipped RGB)
