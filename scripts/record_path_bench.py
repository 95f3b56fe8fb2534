"""Host record path throughput (SURVEY 8(a) rows A-2 .. A-4) on replay-sized records, no GPU needed:
TFRecord indexing + CRC-32C, tf.Example wire parsing, JPEG decoding (host decoder threads) and the Huffman stage of the
split decoder.  Records follow SURVEY 8(d): 512x640 JPEG (quality 90, 4:2:0) of box-filtered noise + the QT-Opt floats.

  python scripts/record_path_bench.py [records]
"""
import io
import os
import sys
import tempfile
import time

import numpy as np
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tensor2robot_b200.utils import dtypes  # noqa: E402
from tensor2robot_b200.utils import example_proto as ep  # noqa: E402
from tensor2robot_b200.utils import jpeg  # noqa: E402
from tensor2robot_b200.utils import tensorspec_utils as utils  # noqa: E402
from tensor2robot_b200.utils import tfdata  # noqa: E402
from tensor2robot_b200.utils import writer  # noqa: E402


def make_records(n):
  out = []
  for i in range(n):
    rng = np.random.default_rng(seed=1234 + i)
    noise = rng.integers(0, 256, (512 + 8, 640 + 8, 3)).astype(np.float32)
    c = np.cumsum(np.cumsum(noise, 0), 1)                       # 8x8 box filter through an integral image
    box = (c[8:, 8:] - c[:-8, 8:] - c[8:, :-8] + c[:-8, :-8]) / 64.0
    img = np.clip((box - 127.5) * 2.0 + 127.5, 0, 255).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format='JPEG', quality=90, subsampling=2)
    f = {'image_1': ep.bytes_feature([buf.getvalue()]), 'world_vector': ep.float_feature(rng.uniform(-1, 1, 3)),
         'vertical_rotation': ep.float_feature(rng.uniform(-1, 1, 2)), 'grasp_success': ep.float_feature([float(rng.random() < 0.3)])}
    for k in ('close_gripper', 'open_gripper', 'terminate_episode', 'gripper_closed'):
      f[k] = ep.float_feature([float(rng.random() < 0.5)])
    f['height_to_bottom'] = ep.float_feature([rng.random()])
    out.append(ep.Example(f))
  return out


def best(fn, reps=3):
  t = 1e30
  for _ in range(reps):
    t0 = time.perf_counter()
    r = fn()
    t = min(t, time.perf_counter() - t0)
  return t, r


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
  tmp = tempfile.mkdtemp()
  w = writer.TFRecordReplayWriter()
  w.open(os.path.join(tmp, 'replay'))
  t0 = time.perf_counter()
  w.write(make_records(n))
  w.close()
  path = os.path.join(tmp, 'replay.tfrecord')
  size = os.path.getsize(path)
  print('%d records, %.1f KB each (generation + JPEG encode + write: %.1f s)' % (n, size / n / 1e3, time.perf_counter() - t0))
  t, f = best(lambda: tfdata.TFRecordFile(path))
  print('index + CRC-32C verify: %.1f ms  (%.2f GB/s, %.0f records/s)' % (t * 1e3, size / t / 1e9, n / t))
  records = list(f)
  tspec = utils.ExtendedTensorSpec
  spec = utils.TensorSpecStruct(image=tspec((512, 640, 3), dtypes.uint8, 'image_1', data_format='jpeg'),
                                world_vector=tspec((3,), dtypes.float32, 'world_vector'),
                                vertical_rotation=tspec((2,), dtypes.float32, 'vertical_rotation'),
                                grasp_success=tspec((1,), dtypes.float32, 'grasp_success'))
  raw_spec = utils.TensorSpecStruct(image=tspec((), dtypes.string, 'image_1'), world_vector=spec.world_vector,
                                    vertical_rotation=spec.vertical_rotation, grasp_success=spec.grasp_success)
  t, raw = best(lambda: tfdata.create_parse_tf_example_fn(raw_spec)(records))
  print('tf.Example wire parse (no image decode): %.2f ms  (%.0f records/s)' % (t * 1e3, n / t))
  t, _ = best(lambda: tfdata.create_parse_tf_example_fn(spec)(records), reps=8)
  print('parse + host JPEG decode (C++ decoder, worker pool): %.1f ms  (%.0f frames/s)' % (t * 1e3, n / t))
  def pil():
    return [np.asarray(Image.open(io.BytesIO(b)).convert('RGB')) for b in jpegs[:16]]
  jpegs = [bytes(b) for b in raw.image]
  t, _ = best(pil, reps=2)
  print('PIL (libjpeg-turbo, decodes under the GIL: threads do not help): %.2f ms / frame  (%.0f frames/s)' % (
      t / 16 * 1e3, 16 / t))
  t, _ = best(lambda: jpeg.entropy_decode(jpegs, pinned=False), reps=8)
  print('split decoder, host half (Huffman threads): %.1f ms  (%.0f frames/s); the IDCT / colour half runs on the GPU' % (
      t * 1e3, n / t))
  print('host threads available: %d' % len(os.sched_getaffinity(0)))


if __name__ == '__main__':
  main()
