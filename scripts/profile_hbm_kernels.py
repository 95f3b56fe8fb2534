"""Launches the HBM-bound kernels of the step at benchmark sizes (batch 512), prints CUDA-event timings and the
algorithmic bytes / achieved GB/s of each, for

  python scripts/profile_hbm_kernels.py                                   # timings (table for profiles/)
  ncu --set full --clock-control none -k regex:'crop_convert|maxpool|optimizer_kernel|jpeg_|bn_' -c 24 \
      -o gpurun_out/r02_hbm_kernels python scripts/profile_hbm_kernels.py

Kernels: crop_convert_distort (vector path), maxpool fwd / bwd (3x3/2 specialisation), momentum / adam / rmsprop,
jpeg_idct + jpeg_color (device half of the split JPEG decoder), bn_apply / bn_bwd_reduce / bn_bwd_apply."""
import ctypes as C
import io
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensor2robot_b200 import _lib, nn   # noqa: E402
from tensor2robot_b200.models import optimizers   # noqa: E402
from tensor2robot_b200.preprocessors import image_ops   # noqa: E402
from tensor2robot_b200.utils import jpeg   # noqa: E402


def timed(fn, reps):
  fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps


def main():
  reps = int(os.environ.get('REPS', '3'))
  dev = torch.device('cuda')
  b = int(os.environ.get('BATCH', '512'))
  rows = []

  def add(name, nbytes, fn):
    ms = timed(fn, reps)
    rows.append({'kernel': name, 'ms': ms, 'algorithmic_gb': nbytes / 1e9, 'gbs': nbytes / ms / 1e6})
    print('%-44s %8.3f ms  %7.2f GB  %7.0f GB/s' % (name, ms, nbytes / 1e9, nbytes / ms / 1e6))

  # crop + convert (+ clip): 3 B read per cropped pixel, 6 B (bf16) written
  frames = torch.randint(0, 256, (b, 512, 640, 3), dtype=torch.uint8, device=dev)
  params = image_ops.identity_params(b, 20, 84)
  add('crop_convert_distort u8->bf16 472x472', b * 472 * 472 * 9,
      lambda: image_ops.crop_convert_distort(frames, (472, 472), params, torch.bfloat16))
  del frames

  # max pool 3x3/2 on the ResNet stem output
  x = torch.randn(b, 236, 236, 64, device=dev).to(torch.bfloat16).requires_grad_(True)
  y = nn.max_pool2d(x, 3, 2, 'SAME')
  dy = torch.randn_like(y)
  nx, ny = x.numel(), y.numel()
  add('maxpool_fwd 3x3/2 236->118 C=64', nx * 2 + ny * 3, lambda: nn.max_pool2d(x, 3, 2, 'SAME'))
  add('maxpool_bwd 3x3/2 236->118 C=64', ny * 3 + nx * 2, lambda: y.backward(dy, retain_graph=True))
  del x, y, dy

  # optimizers over the ResNet-50 critic's 23.9 M parameters (+EMA, +bf16 copy)
  vs = nn.VariableStore('cuda')
  with nn.variable_store(vs):
    vs.get_variable('w', (23_900_000 // 64 * 64,), lambda s, r: np.zeros(s, np.float32), regularize=True)
    vs.finalize()
  n = vs.flat.numel()
  for name, opt, per in (('momentum', optimizers.MomentumOptimizer(1e-4, 0.9), 4 * 3 + 4 * 2 + 8 + 2),
                         ('adam', optimizers.AdamOptimizer(1e-4), 4 * 4 + 4 * 3 + 8 + 2),
                         ('rmsprop', optimizers.RMSPropOptimizer(1e-4, momentum=0.9, epsilon=1.0), 4 * 4 + 4 * 3 + 8 + 2)):
    wrapped = optimizers.MovingAverageOptimizer(opt, 0.9999)
    add('%s + l2 + EMA + bf16 refresh, 23.9 M params' % name, n * per, lambda w=wrapped: w.apply_gradients(vs, 0, 1.0))

  # batch-norm passes on the largest activation (256 channels at 118x118)
  r, c = b * 118 * 118, 256
  xb = torch.randn(r, c, device=dev).to(torch.bfloat16)
  yb = torch.empty_like(xb)
  dyb = torch.randn(r, c, device=dev).to(torch.bfloat16)
  sc, sh = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev)
  mean, inv = torch.randn(c, device=dev), torch.rand(c, device=dev) + 0.5
  red = torch.zeros(2 * c, dtype=torch.float64, device=dev)
  dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
  st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
  p = lambda t: C.c_void_p(t.data_ptr())
  add('bn_apply_rows 512x118x118x256', r * c * 4,
      lambda: _lib.call('t2r_bn_apply', p(xb), p(yb), r, c, p(sc), p(sh), None, 118 * 118, 1, st))
  add('bn_bwd (reduce + apply) 512x118x118x256', r * c * 10,
      lambda: _lib.call('t2r_bn_backward', p(dyb), p(xb), None, p(yb), r, c, None, p(mean), p(inv), p(sc), p(sh), 1, p(red),
                        p(dg), p(db), st))
  del xb, yb, dyb

  # device half of the split JPEG decoder on 4:2:0 512x640 frames
  from PIL import Image
  rng = np.random.default_rng(0)
  jpegs = []
  for i in range(8):
    noise = rng.integers(0, 256, (520, 648, 3)).astype(np.float32)
    cs = np.cumsum(np.cumsum(noise, 0), 1)
    box = (cs[8:, 8:] - cs[:-8, 8:] - cs[8:, :-8] + cs[:-8, :-8]) / 64.0
    buf = io.BytesIO()
    Image.fromarray(np.clip((box - 127.5) * 2 + 127.5, 0, 255).astype(np.uint8)).save(buf, format='JPEG', quality=90, subsampling=2)
    jpegs.append(buf.getvalue())
  nb = min(b, 256)
  geom, coef, qt = jpeg.entropy_decode([jpegs[i % 8] for i in range(nb)])
  coef_d, qt_d = coef.to(dev), qt.to(dev)
  planes = torch.empty((nb, int(geom.coef_count)), dtype=torch.uint8, device=dev)
  out = torch.empty((nb, 512, 640, 3), dtype=torch.uint8, device=dev)
  add('jpeg_idct + jpeg_color 4:2:0 512x640 (%d frames)' % nb, nb * (int(geom.coef_count) * (2 + 1 + 1) + 512 * 640 * 3),
      lambda: _lib.call('t2r_jpeg_idct_color', p(coef_d), p(qt_d), C.byref(geom), p(planes), p(out), nb, int(geom.coef_count), 3, st))
  path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'r02_hbm_kernels.json')
  try:
    with open(path, 'w') as f:
      json.dump(rows, f, indent=1)
  except OSError:
    pass


if __name__ == '__main__':
  main()
