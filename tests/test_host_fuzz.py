"""Robustness of the host-side parsers of untrusted bytes (csrc/jpeg_host.cc, csrc/host_io.cc): mutated JPEG streams
and records are either decoded or rejected with an error status - never a crash.  The python-level loop runs against the
shipped library; the AddressSanitizer / UBSan harness (tests/native/fuzz/host_fuzz.cc, scripts/host_fuzz.sh) rebuilds the
two sources with instrumentation so silent out-of-bounds accesses abort as well."""
import io
import os
import shutil
import subprocess

import numpy as np
import pytest
from PIL import Image

from tensor2robot_b200 import _lib
from tensor2robot_b200.utils import jpeg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _seeds():
  rng = np.random.RandomState(0)
  a = rng.randint(0, 256, (48, 64, 3)).astype(np.uint8)
  out = []
  for kw in (dict(subsampling=0), dict(subsampling=1), dict(subsampling=2), dict(subsampling=2, restart_marker_blocks=2)):
    buf = io.BytesIO()
    Image.fromarray(a).save(buf, format='JPEG', quality=80, **kw)
    out.append(buf.getvalue())
  buf = io.BytesIO()
  Image.fromarray(a[..., 0]).save(buf, format='JPEG')
  out.append(buf.getvalue())
  return out


def test_mutated_jpegs_never_crash_the_decoder():
  rng = np.random.RandomState(1)
  seeds = _seeds()
  accepted = rejected = 0
  for it in range(3000):
    b = bytearray(seeds[it % len(seeds)])
    mode = rng.randint(4)
    if mode == 0:
      for _ in range(rng.randint(1, 6)):
        b[rng.randint(len(b))] = rng.randint(256)
    elif mode == 1:
      b = b[:rng.randint(2, len(b))]
    elif mode == 2:
      i = rng.randint(len(b))
      b[i:i] = bytes(rng.randint(0, 256, rng.randint(1, 20)).astype(np.uint8))
    else:
      b[rng.randint(2, min(len(b), 700))] = 0xFF
    try:
      jpeg.entropy_decode([bytes(b)], pinned=False)
      accepted += 1
    except (jpeg.UnsupportedJpeg, _lib.T2RError):
      rejected += 1
  assert accepted > 100 and rejected > 100


def test_known_bad_headers_are_rejected():
  seed = _seeds()[2]
  sof = seed.index(b'\xff\xc0')
  huge = bytearray(seed)
  huge[sof + 5:sof + 9] = b'\xff\xff\xff\xff'                    # 65535 x 65535 frame
  with pytest.raises(jpeg.UnsupportedJpeg, match='exceeds the supported size'):
    jpeg.parse(bytes(huge))
  sos = seed.index(b'\xff\xda')
  bad_table = bytearray(seed)
  bad_table[sos + 6] = 0xF0                                     # DC table selector 15 for the first component
  with pytest.raises(jpeg.UnsupportedJpeg, match='Huffman table'):
    jpeg.entropy_decode([bytes(bad_table)], pinned=False)
  dht = seed.index(b'\xff\xc4')
  oversubscribed = bytearray(seed)
  oversubscribed[dht + 5:dht + 8] = b'\x05\x05\x05'             # 5 codes of length 1: impossible
  with pytest.raises(jpeg.UnsupportedJpeg):
    jpeg.entropy_decode([bytes(oversubscribed)], pinned=False)
  wrong_component = bytearray(seed)
  wrong_component[sos + 5] = 0x7B                               # the scan names a component id the frame does not have
  with pytest.raises(jpeg.UnsupportedJpeg, match='Huffman table'):
    jpeg.entropy_decode([bytes(wrong_component)], pinned=False)


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++ with libasan')
def test_sanitizer_harness():
  probe = subprocess.run(['g++', '-fsanitize=address,undefined', '-x', 'c++', '-', '-o', '/dev/null'], input=b'int main(){}',
                         capture_output=True)
  if probe.returncode != 0:
    pytest.skip('this toolchain cannot link the sanitizer runtimes')
  out = subprocess.run([os.path.join(ROOT, 'scripts', 'host_fuzz.sh'), '6000'], capture_output=True, text=True, timeout=600)
  assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
  assert 'no memory error' in out.stdout


def test_mutated_records_only_raise_value_errors():
  """The python parse layer (C++ wire parser + image decoding) turns every malformed record - truncated protos, corrupt
  JPEG payloads, wrong counts - into a ValueError, the analogue of TF's InvalidArgumentError."""
  from oracle import tfrecord
  from tensor2robot_b200.utils import dtypes
  from tensor2robot_b200.utils import tensorspec_utils as utils
  from tensor2robot_b200.utils import tfdata
  tspec = utils.ExtendedTensorSpec
  feature_spec = utils.TensorSpecStruct(
      state=utils.TensorSpecStruct(image=tspec((64, 64, 3), dtypes.uint8, 'state/image', data_format='jpeg')),
      action=utils.TensorSpecStruct(pose=tspec((2,), dtypes.float32, 'pose')))
  label_spec = utils.TensorSpecStruct(reward=tspec((1,), dtypes.float32, 'reward'))
  parse = tfdata.create_parse_tf_example_fn(feature_spec, label_spec)
  records = tfrecord.read_tfrecords(os.path.join(ROOT, 'tests', 'golden', 'pose_env_test_data.tfrecord'))[:8]
  rng = np.random.RandomState(0)
  parsed = rejected = 0
  for it in range(800):
    b = bytearray(records[it % 8])
    mode = rng.randint(3)
    if mode == 0:
      for _ in range(rng.randint(1, 6)):
        b[rng.randint(len(b))] = rng.randint(256)
    elif mode == 1:
      b = b[:rng.randint(1, len(b))]
    else:
      i = rng.randint(len(b))
      b[i:i] = bytes(rng.randint(0, 256, rng.randint(1, 9)).astype(np.uint8))
    try:
      parse([bytes(b), records[0]])
      parsed += 1
    except ValueError:
      rejected += 1
  assert parsed > 50 and rejected > 50
