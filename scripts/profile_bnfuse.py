"""Launches the batch-norm-fused convolution kernels of ResNet-50 layer 1 / 2 at batch 512 a few times, for ncu:

  ncu --set full --clock-control none --import-source on -k regex:conv_ -c 12 -o gpurun_out/bnfuse python scripts/profile_bnfuse.py

Prints CUDA-event timings of each variant when run without a profiler."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tensor2robot_b200 import _lib, nn   # noqa: E402


def main():
  reps = int(os.environ.get('REPS', '1'))
  n, hw, cin, cout = 512, 118, 256, 64
  dev = torch.device('cuda')
  x = (torch.randn(n, hw, hw, cin, device=dev) * 0.7).to(torch.bfloat16)
  dy = torch.randn(n, hw, hw, cout, device=dev).to(torch.bfloat16)
  w = (torch.randn(cout, 1, 1, cin, device=dev) / 16).float()
  wb = w.to(torch.bfloat16)
  wd = torch.empty(cin * cout, dtype=torch.bfloat16, device=dev)
  st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
  p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
  _lib.call('t2r_pack_weights', p(w), None, p(wd), cout, 1, cin, st)
  scale = torch.rand(cin, device=dev) + 0.5
  shift = torch.randn(cin, device=dev) * 0.3
  d = nn._conv_desc(n, hw, hw, cin, cout, 1, 1, 1, 0, 0, hw, hw, 0)
  y = torch.empty(n, hw, hw, cout, dtype=torch.bfloat16, device=dev)
  g = torch.empty_like(x)
  dw = torch.zeros(cout, 1, 1, cin, device=dev)
  red = torch.zeros(2 * cin, dtype=torch.float64, device=dev)
  stats = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
  variants = [
      ('fprop plain', lambda: _lib.call('t2r_conv2d_fprop_stats', C.byref(d), p(x), p(wb), None, None, p(y), p(stats), st)),
      ('fprop bnrelu', lambda: _lib.call('t2r_conv2d_fprop_bnrelu', C.byref(d), p(x), p(scale), p(shift), p(wb), None, p(y), p(stats), st)),
      ('wgrad plain', lambda: _lib.call('t2r_conv2d_wgrad', C.byref(d), p(x), p(dy), p(dw), st)),
      ('wgrad bnrelu', lambda: _lib.call('t2r_conv2d_wgrad_bnrelu', C.byref(d), p(x), p(scale), p(shift), p(dy), p(dw), st)),
      ('dgrad plain', lambda: _lib.call('t2r_conv2d_dgrad', C.byref(d), p(dy), p(wd), p(g), 0, st)),
      ('dgrad bnrelu', lambda: _lib.call('t2r_conv2d_dgrad_bnrelu', C.byref(d), p(dy), p(wd), p(x), p(scale), p(shift), p(g), 0, p(red), st)),
  ]
  for name, fn in variants:
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
      fn()
    e1.record()
    torch.cuda.synchronize()
    print('%-14s %.3f ms' % (name, e0.elapsed_time(e1) / reps))


if __name__ == '__main__':
  main()
