"""Throughput of the split JPEG decoder on replay-frame sized images (512x640, 4:2:0, q90)."""
import io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
from tensor2robot_b200.utils import jpeg
rng = np.random.RandomState(0)
yy, xx = np.mgrid[0:512, 0:640].astype(np.float32)
imgs = []
for i in range(16):
  a = np.stack([127 + 100 * np.sin(xx / (9 + i)), 127 + 100 * np.cos(yy / 7), (xx + yy + 13 * i) % 256], -1)
  a = np.clip(a + rng.uniform(-20, 20, a.shape), 0, 255).astype(np.uint8)
  buf = io.BytesIO(); Image.fromarray(a).save(buf, format='JPEG', quality=90, subsampling=2); imgs.append(buf.getvalue())
batch = [imgs[i % 16] for i in range(512)]
print('bytes per image', np.mean([len(b) for b in batch]))
for _ in range(2):
  t0 = time.perf_counter(); geom, coef, qt = jpeg.entropy_decode(batch); t1 = time.perf_counter()
  out = jpeg.decode_batch(batch); torch.cuda.synchronize(); t2 = time.perf_counter()
print('host Huffman: %.1f ms / 512 images (%.0f img/s); Huffman + H2D + device: %.1f ms (%.0f img/s)' % (
    (t1 - t0) * 1e3, 512 / (t1 - t0), (t2 - t1) * 1e3, 512 / (t2 - t1)))
t0 = time.perf_counter()
for b in batch[:64]:
  np.asarray(Image.open(io.BytesIO(b)).convert('RGB'))
t1 = time.perf_counter()
print('PIL (libjpeg-turbo) single thread: %.2f ms / image' % ((t1 - t0) / 64 * 1e3))
