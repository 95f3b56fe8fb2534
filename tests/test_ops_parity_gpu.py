"""Per-op GPU parity through the C-ABI against the oracle ops, on identical inputs: the tight
correctness gates (the end-to-end network tests are conditioning-limited).

Tolerances: operands are bf16-exact on both sides, accumulation is fp32 on the GPU and fp32/fp64 in
the oracle, outputs are stored as bf16 (relative rounding 2^-9 = 2.0e-3): conv/BN outputs are
asserted within 4e-3 relative of the output scale; fp32 outputs (wgrad, statistics) within 1e-4."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BF16_TOL = 4e-3


def _p(t):
  return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _bf16_exact(rng, shape, scale=1.0):
  return torch.from_numpy((rng.standard_normal(shape) * scale).astype(np.float32)).to(torch.bfloat16).float()


def _check(name, got, ref, rel, scale=None):
  got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
  scale = scale if scale is not None else max(np.abs(ref).max(), 1e-12)
  err = np.abs(got - ref).max() / scale
  print('%-40s max err / scale = %.3e' % (name, err))
  assert err < rel, (name, err)


CONV_CASES = [
    # n, h, w, cin, cout, k, stride, padding
    (2, 27, 27, 64, 64, 3, 1, 'SAME'),        # Grasping44 conv8..13
    (1, 40, 37, 64, 64, 5, 1, 'SAME'),        # Grasping44 conv2..7 geometry (odd sizes)
    (2, 14, 14, 64, 64, 3, 1, 'VALID'),       # conv14..16
    (2, 30, 30, 128, 128, 3, 2, 'FIXED'),     # ResNet strided 3x3 (fixed_padding + VALID)
    (2, 59, 59, 256, 512, 1, 2, 'FIXED'),     # ResNet projection shortcut
    (2, 15, 15, 512, 2048, 1, 1, 'SAME'),     # ResNet bottleneck expansion
    (1, 1, 70, 4096, 64, 1, 1, 'VALID'),      # fc0 as a GEMM
]


@pytest.mark.parametrize('case', CONV_CASES)
def test_conv_fprop_dgrad_wgrad_match_oracle(case):
  from oracle import tf_ops
  from tensor2robot_b200 import _lib, nn
  n, h, w, cin, cout, k, stride, padding = case
  rng = np.random.RandomState(hash(case) % 2**31)
  x = _bf16_exact(rng, (n, h, w, cin))
  w_hwio = _bf16_exact(rng, (k, k, cin, cout), 1.0 / np.sqrt(k * k * cin))
  ho, wo, pt, pl = nn.conv_geometry(h, w, k, k, stride, padding)
  dy = _bf16_exact(rng, (n, ho, wo, cout))
  # oracle (fp32 autograd)
  xo, wo_ = x.clone().requires_grad_(True), w_hwio.clone().requires_grad_(True)
  yo = tf_ops.conv2d_fixed_padding(xo, wo_, stride) if padding == 'FIXED' else tf_ops.conv2d(xo, wo_, stride, padding)
  assert tuple(yo.shape) == (n, ho, wo, cout)
  yo.backward(dy)
  # engine through the C-ABI
  d = nn._conv_desc(n, h, w, cin, cout, k, k, stride, pt, pl, ho, wo)
  w_ohwi = w_hwio.permute(3, 0, 1, 2).contiguous().cuda()
  wf = torch.empty(w_ohwi.shape, dtype=torch.bfloat16, device='cuda')
  wd = torch.empty(w_ohwi.numel(), dtype=torch.bfloat16, device='cuda')
  _lib.call('t2r_pack_weights', _p(w_ohwi), _p(wf), _p(wd), cout, k * k, cin, None)
  xg, dyg = x.cuda().to(torch.bfloat16), dy.cuda().to(torch.bfloat16)
  y = torch.empty((n, ho, wo, cout), dtype=torch.bfloat16, device='cuda')
  dx = torch.empty_like(xg)
  dw = torch.zeros(w_ohwi.shape, dtype=torch.float32, device='cuda')
  _lib.call('t2r_conv2d_fprop', C.byref(d), _p(xg), _p(wf), None, None, _p(y), None)
  _lib.call('t2r_conv2d_dgrad', C.byref(d), _p(dyg), _p(wd), _p(dx), 0, None)
  _lib.call('t2r_conv2d_wgrad', C.byref(d), _p(xg), _p(dyg), _p(dw), None)
  torch.cuda.synchronize()
  _check('fprop %s' % (case,), y.float().cpu(), yo.detach(), BF16_TOL)
  _check('dgrad %s' % (case,), dx.float().cpu(), xo.grad, BF16_TOL)
  _check('wgrad %s' % (case,), dw.cpu().permute(1, 2, 3, 0), wo_.grad, 2e-4)


def test_stem_im2col_conv_matches_oracle():
  """6x6/2 SAME on 3 channels (Grasping44 conv1_1) and 7x7/2 fixed padding (ResNet stem)."""
  from oracle import tf_ops
  from tensor2robot_b200 import nn
  rng = np.random.RandomState(0)
  for k, padding in ((6, 'SAME'), (7, 'FIXED')):
    x = _bf16_exact(rng, (2, 61, 53, 3)).abs()
    w_hwio = _bf16_exact(rng, (k, k, 3, 64), 0.1)
    bias = torch.from_numpy(rng.standard_normal(64).astype(np.float32))
    xo, wo_ = x.clone(), w_hwio.clone().requires_grad_(True)
    yo = (tf_ops.conv2d_fixed_padding(xo, wo_, 2) if padding == 'FIXED' else tf_ops.conv2d(xo, wo_, 2, padding)) + bias
    dy = _bf16_exact(rng, tuple(yo.shape))
    yo.backward(dy)
    vs = nn.VariableStore('cuda')
    with nn.variable_store(vs):
      xg = x.cuda().to(torch.bfloat16)
      y = nn.conv2d(xg, 64, k, 2, padding, use_bias=True, scope='stem')
      vs.finalize()
      vs.import_tf({'stem/weights': w_hwio.numpy(), 'stem/biases': bias.numpy()})
      y = nn.conv2d(xg, 64, k, 2, padding, use_bias=True, scope='stem')
      vs.zero_grad()
      y.backward(dy.cuda().to(torch.bfloat16))
    torch.cuda.synchronize()
    _check('stem k%d fprop' % k, y.detach().float().cpu(), yo.detach(), BF16_TOL)
    grads = vs.export_tf_grads()
    _check('stem k%d wgrad' % k, grads['stem/weights'], wo_.grad, 2e-4)
    _check('stem k%d bias grad' % k, grads['stem/biases'], dy.sum((0, 1, 2)), 2e-4)


@pytest.mark.parametrize('rows,c,scale,relu', [(5000, 64, True, True), (333, 256, False, True), (64, 2048, True, False)])
def test_batch_norm_forward_backward_match_oracle(rows, c, scale, relu):
  from oracle import tf_ops
  from tensor2robot_b200 import nn
  rng = np.random.RandomState(rows + c)
  x = _bf16_exact(rng, (rows, c), 2.0) + _bf16_exact(rng, (1, c))
  x = x.to(torch.bfloat16).float()
  dy = _bf16_exact(rng, (rows, c))
  variables = {'bn/beta': torch.from_numpy(rng.standard_normal(c).astype(np.float32) * 0.3).requires_grad_(True),
               'bn/moving_mean': torch.from_numpy(rng.standard_normal(c).astype(np.float32)),
               'bn/moving_variance': torch.from_numpy(rng.uniform(0.5, 2, c).astype(np.float32))}
  if scale:
    variables['bn/gamma'] = torch.from_numpy(1 + 0.3 * rng.standard_normal(c).astype(np.float32)).requires_grad_(True)
  xo = x.clone().requires_grad_(True)
  updates = {}
  yo = tf_ops.batch_norm(xo, variables, 'bn', True, 0.997, 1e-3, scale, updates)
  if relu:
    yo = torch.relu(yo)
  yo.backward(dy)
  vs = nn.VariableStore('cuda')
  with nn.variable_store(vs):
    xg = x.cuda().to(torch.bfloat16).requires_grad_(True)
    nn.batch_norm(xg.detach(), False, scope='bn', scale=scale, relu=relu, momentum=0.997, eps=1e-3)
    vs.finalize()
    vs.import_tf({k: v.detach().numpy() for k, v in variables.items()})
    y = nn.batch_norm(xg, True, scope='bn', scale=scale, relu=relu, momentum=0.997, eps=1e-3)
    vs.zero_grad()
    y.backward(dy.cuda().to(torch.bfloat16))
  torch.cuda.synchronize()
  new = vs.export_tf()
  grads = vs.export_tf_grads()
  _check('bn y', y.detach().float().cpu(), yo.detach(), BF16_TOL)
  _check('bn dx', xg.grad.float().cpu(), xo.grad, BF16_TOL)
  _check('bn dbeta', grads['bn/beta'], variables['bn/beta'].grad, 2e-4)
  if scale:
    _check('bn dgamma', grads['bn/gamma'], variables['bn/gamma'].grad, 2e-4)
  _check('bn moving_mean', new['bn/moving_mean'], updates['bn/moving_mean'], 1e-5)
  _check('bn moving_variance', new['bn/moving_variance'], updates['bn/moving_variance'], 1e-5)
  # inference mode reads the moving statistics
  with torch.no_grad(), nn.variable_store(vs):
    yi = nn.batch_norm(xg.detach(), False, scope='bn', scale=scale, relu=relu, momentum=0.997, eps=1e-3)
  ev = {k: (torch.from_numpy(new[k]) if k in new else v.detach()) for k, v in variables.items()}
  yio = tf_ops.batch_norm(x, ev, 'bn', False, 0.997, 1e-3, scale)
  _check('bn inference', yi.float().cpu(), torch.relu(yio) if relu else yio, BF16_TOL)


@pytest.mark.parametrize('h,w,k,s,padding', [(236, 236, 3, 3, 'SAME'), (79, 79, 3, 3, 'SAME'), (27, 27, 2, 2, 'SAME'),
                                             (47, 46, 3, 2, 'SAME')])
def test_max_pool_matches_oracle(h, w, k, s, padding):
  from oracle import tf_ops
  from tensor2robot_b200 import nn
  rng = np.random.RandomState(h)
  x = _bf16_exact(rng, (2, h, w, 64))
  xo = x.clone().requires_grad_(True)
  yo = tf_ops.max_pool(xo, k, s, padding)
  dy = _bf16_exact(rng, tuple(yo.shape))
  yo.backward(dy)
  xg = x.cuda().to(torch.bfloat16).requires_grad_(True)
  y = nn.max_pool2d(xg, k, s, padding)
  y.backward(dy.cuda().to(torch.bfloat16))
  torch.cuda.synchronize()
  assert torch.equal(y.detach().float().cpu(), yo.detach())            # selection: exact
  _check('maxpool dx', xg.grad.float().cpu(), xo.grad, BF16_TOL)


def test_crop_convert_distort_matches_oracle():
  from oracle import image_ops as oracle
  from tensor2robot_b200.preprocessors import image_ops
  rng = np.random.RandomState(3)
  frames = rng.randint(0, 256, (3, 96, 120, 3)).astype(np.uint8)
  frames[0, :, :, 0] = np.arange(120)[None, :]                  # ramp image: crop offsets are checkable
  frames[0, :, :, 1] = np.arange(96)[:, None]                   # (image_transformations_test.py:133-161)
  d = torch.from_numpy(frames).cuda()
  # crop + convert only: bit exact with x * (1/255)
  p = image_ops.identity_params(3, 5, 11)
  out = image_ops.crop_convert_distort(d, (80, 100), p, torch.float32)
  ref = oracle.convert_image_dtype_f32(oracle.crop(frames, 5, 11, 80, 100))
  assert np.array_equal(out.cpu().numpy(), ref)
  assert out[0, 0, 0, 0].item() == np.float32(11) * np.float32(1 / 255.0) and out[0, 0, 0, 1].item() == np.float32(5) * np.float32(1 / 255.0)
  # centre crop geometry of the QT-Opt preprocessor: (512, 640) -> (472, 472) => offsets (20, 84)
  assert ((512 - 472) // 2, (640 - 472) // 2) == (20, 84)
  # bf16 output = round-to-nearest-even of the float result
  out_b = image_ops.crop_convert_distort(d, (80, 100), p, torch.bfloat16)
  assert torch.equal(out_b.cpu(), torch.from_numpy(ref).to(torch.bfloat16))
  # every photometric op, alone and chained (fp32; HSV round trip tolerance 1e-5, SURVEY A-11)
  for kwargs in ({'brightness_delta': 0.1}, {'saturation_scale': 0.6}, {'saturation_scale': 1.4}, {'hue_delta': 0.15},
                 {'hue_delta': -0.2}, {'contrast_scale': 1.3},
                 {'brightness_delta': -0.05, 'saturation_scale': 1.2, 'hue_delta': 0.1, 'contrast_scale': 0.7}):
    p = image_ops.identity_params(3, 5, 11)
    for k_, v_ in kwargs.items():
      p[k_] = v_
    out = image_ops.crop_convert_distort(d, (80, 100), p, torch.float32).cpu().numpy()
    want = oracle.distort(ref, **kwargs)
    err = np.abs(out - want).max()
    print('distort %-70s max abs err %.2e' % (kwargs, err))
    assert err < 2e-5, kwargs
  # noise: N(0, sigma) added before the clip, deterministic in (seed, offset)
  p = image_ops.identity_params(3, 5, 11)
  p['noise_stddev'] = 0.05
  a = image_ops.crop_convert_distort(d, (80, 100), p, torch.float32, seed=9, offset=1).cpu().numpy()
  b = image_ops.crop_convert_distort(d, (80, 100), p, torch.float32, seed=9, offset=1).cpu().numpy()
  assert np.array_equal(a, b)
  inner = (ref > 0.2) & (ref < 0.8)
  assert abs((a - ref)[inner].std() - 0.05) < 0.003 and abs((a - ref)[inner].mean()) < 0.002


def test_resize_bilinear_legacy_matches_oracle():
  from oracle import image_ops as oracle
  from tensor2robot_b200.preprocessors import image_ops
  rng = np.random.RandomState(4)
  x = rng.uniform(0, 1, (2, 45, 45, 3)).astype(np.float32)
  for out_hw in ((20, 20), (15, 25), (90, 60)):
    got = image_ops.resize_bilinear_legacy(torch.from_numpy(x).cuda(), out_hw).cpu().numpy()
    want = oracle.resize_bilinear_legacy(x, *out_hw)
    assert np.abs(got - want).max() < 1e-6, out_hw


def test_optimizers_match_tf_formulas():
  from tensor2robot_b200 import nn
  from tensor2robot_b200.models import optimizers
  rng = np.random.RandomState(5)
  for make, ref_fn in ((lambda: optimizers.MomentumOptimizer(0.1, 0.9), 'momentum'),
                       (lambda: optimizers.AdamOptimizer(0.01), 'adam'),
                       (lambda: optimizers.RMSPropOptimizer(0.05, decay=0.9, momentum=0.9, epsilon=1.0), 'rmsprop')):
    vs = nn.VariableStore('cuda')
    w0 = rng.standard_normal((7, 13)).astype(np.float32)
    with nn.variable_store(vs):
      v = vs.get_variable('w', (7, 13), lambda s, r: w0, regularize=True)
      vs.finalize()
    opt = optimizers.MovingAverageOptimizer(make(), 0.99)
    opt.l2_regularization = 0.01
    w = w0.astype(np.float64).copy()
    ema = w.copy()
    m = np.zeros_like(w); vv = np.zeros_like(w)
    rms = np.ones_like(w)              # tf.train.RMSPropOptimizer initialises its rms slot to one
    for step in range(4):
      g = rng.standard_normal((7, 13)).astype(np.float32)
      v.grad.copy_(torch.from_numpy(g))
      opt.apply_gradients(vs, step, grad_scale=0.5)
      gg = 0.5 * g + 0.01 * w
      if ref_fn == 'momentum':
        m = 0.9 * m + gg
        w = w - 0.1 * m
      elif ref_fn == 'rmsprop':           # research/qtopt/optimizer_builder.py:76-81 (decay .9, momentum, epsilon 1.0)
        rms = 0.9 * rms + 0.1 * gg * gg
        m = 0.9 * m + 0.05 * gg / np.sqrt(rms + 1.0)
        w = w - m
      else:
        t = step + 1
        m = 0.9 * m + 0.1 * gg
        vv = 0.999 * vv + 0.001 * gg * gg
        w = w - 0.01 * np.sqrt(1 - 0.999**t) / (1 - 0.9**t) * m / (np.sqrt(vv) + 1e-8)
      ema = 0.99 * ema + 0.01 * w
    torch.cuda.synchronize()
    np.testing.assert_allclose(v.data.cpu().numpy(), w, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(opt.shadow(vs)[:v.numel].view(7, 13).cpu().numpy(), ema, rtol=2e-5, atol=2e-6)
    assert torch.equal(v.bf16.float().cpu(), v.data.to(torch.bfloat16).float().cpu())


@pytest.mark.gpu
@pytest.mark.parametrize('relu', [True, False])
def test_film_batch_norm_forward_backward_match_oracle(relu):
  """FiLM-conditioned batch norm (layers/film_resnet_model.py:108-115 on top of :50-57): y, dx, dfilm,
  dgamma, dbeta against torch autograd on the oracle's formula."""
  from oracle import tf_ops
  from tensor2robot_b200 import nn
  n, hw, c = 6, 35, 64
  rng = np.random.RandomState(7)
  x = (_bf16_exact(rng, (n, hw, 1, c), 2.0) + _bf16_exact(rng, (1, 1, 1, c))).to(torch.bfloat16).float()
  dy = _bf16_exact(rng, (n, hw, 1, c))
  film = torch.from_numpy(rng.standard_normal((n, 2 * c)).astype(np.float32) * 0.5)
  variables = {'bn/beta': torch.from_numpy(rng.standard_normal(c).astype(np.float32) * 0.3).requires_grad_(True),
               'bn/gamma': torch.from_numpy(1 + 0.3 * rng.standard_normal(c).astype(np.float32)).requires_grad_(True),
               'bn/moving_mean': torch.zeros(c), 'bn/moving_variance': torch.ones(c)}
  xo, fo = x.clone().requires_grad_(True), film.clone().requires_grad_(True)
  yo = tf_ops.batch_norm(xo, variables, 'bn', True, 0.997, 1e-5, True, {})
  yo = (1 + fo[:, None, None, :c]) * yo + fo[:, None, None, c:]
  if relu:
    yo = torch.relu(yo)
  yo.backward(dy)
  vs = nn.VariableStore('cuda')
  with nn.variable_store(vs):
    xg = x.cuda().to(torch.bfloat16).requires_grad_(True)
    fg = film.cuda().requires_grad_(True)
    nn.batch_norm(xg.detach(), False, scope='bn', scale=True, relu=relu, momentum=0.997, eps=1e-5, film=fg.detach())
    vs.finalize()
    vs.import_tf({k: v.detach().numpy() for k, v in variables.items()})
    y = nn.batch_norm(xg, True, scope='bn', scale=True, relu=relu, momentum=0.997, eps=1e-5, film=fg)
    vs.zero_grad()
    y.backward(dy.cuda().to(torch.bfloat16))
  torch.cuda.synchronize()
  grads = vs.export_tf_grads()
  _check('film bn y', y.detach().float().cpu(), yo.detach(), BF16_TOL)
  _check('film bn dx', xg.grad.float().cpu(), xo.grad, BF16_TOL)
  _check('film bn dfilm', fg.grad.cpu(), fo.grad, 2e-4)
  _check('film bn dgamma', grads['bn/gamma'], variables['bn/gamma'].grad, 2e-4)
  _check('film bn dbeta', grads['bn/beta'], variables['bn/beta'].grad, 2e-4)


@pytest.mark.gpu
@pytest.mark.parametrize('n,h,w,c', [(3, 21, 17, 32), (2, 6, 6, 64), (1, 50, 64, 8)])
def test_spatial_softmax_matches_oracle(n, h, w, c):
  """layers/spatial_softmax.py:29-88: expected feature points (interleaved x, y), heat map, and the
  gradient of a linear functional of the points against torch autograd on the same formula."""
  from oracle import spatial_softmax as oracle
  from tensor2robot_b200 import nn
  rng = np.random.RandomState(n * 100 + h)
  x = _bf16_exact(rng, (n, h, w, c), 3.0)
  pts_o, heat_o = oracle.build_spatial_softmax(x.numpy())
  xg = x.cuda().to(torch.bfloat16).requires_grad_(True)
  pts, heat = nn.spatial_softmax(xg, return_softmax=True)
  _check('spatial softmax points', pts.detach().cpu(), torch.from_numpy(pts_o), 2e-5)
  _check('spatial softmax map', heat.float().cpu(), torch.from_numpy(heat_o), BF16_TOL)
  # layout: x of channel k at 2k, y at 2k+1; a peak at the last column / first row gives (+1, -1)
  peak = torch.full((1, h, w, 8), -30.0)
  peak[0, 0, w - 1, :] = 30.0
  p = nn.spatial_softmax(peak.cuda().to(torch.bfloat16)).cpu()
  assert torch.allclose(p[0, 0::2], torch.ones(8), atol=1e-4) and torch.allclose(p[0, 1::2], -torch.ones(8), atol=1e-4)
  # backward
  wgt = torch.from_numpy(rng.standard_normal((n, 2 * c)).astype(np.float32))
  (pts * wgt.cuda()).sum().backward()
  xo = x.clone().requires_grad_(True)
  f = xo.permute(0, 3, 1, 2).reshape(-1, h * w)
  s = torch.softmax(f, 1)
  jj, ii = torch.meshgrid(torch.arange(w), torch.arange(h), indexing='xy')
  xp = (2.0 * jj.reshape(-1) / (w - 1.0) - 1.0).float()
  yp = (2.0 * ii.reshape(-1) / (h - 1.0) - 1.0).float()
  po = torch.cat([(s * xp).sum(1, keepdim=True), (s * yp).sum(1, keepdim=True)], 1).reshape(-1, 2 * c)
  (po * wgt).sum().backward()
  _check('spatial softmax dx', xg.grad.float().cpu(), xo.grad, BF16_TOL)


@pytest.mark.gpu
def test_bcz_preprocess_distorts_after_resize():
  """preprocessors/distortion.py:56-107 (BC-Z): uint8 -> float, crop, TF1-legacy bilinear resize, THEN the
  photometric distortions on the float image (t2r_distort_f32) and the clip."""
  from oracle import image_ops as oracle
  from tensor2robot_b200.preprocessors import distortion, image_ops, image_transformations
  rng = np.random.RandomState(11)
  frames = rng.randint(0, 256, (3, 96, 120, 3)).astype(np.uint8)
  d = torch.from_numpy(frames).cuda()
  # float-image distortion kernel against the oracle, every op chained
  x = rng.uniform(0, 1, (3, 40, 50, 3)).astype(np.float32)
  kwargs = {'brightness_delta': -0.05, 'saturation_scale': 1.2, 'hue_delta': 0.1, 'contrast_scale': 0.7}
  p = image_ops.identity_params(3)
  for k_, v_ in kwargs.items():
    p[k_] = v_
  got = image_ops.distort_f32(torch.from_numpy(x).cuda(), p).cpu().numpy()
  assert np.abs(got - oracle.distort(x, **kwargs)).max() < 2e-5
  # the whole BC-Z chain in eval mode (centre crop, no distortion) and in train mode with fixed draws
  out = distortion.preprocess_image(d, 'eval', False, (96, 120), (30, 40), crop_size=(80, 100)).cpu().numpy()
  ref = oracle.resize_bilinear_legacy(oracle.convert_image_dtype_f32(oracle.crop(frames, 8, 10, 80, 100)), 30, 40)
  assert np.abs(out - np.clip(ref, 0, 1)).max() < 1e-6
  image_transformations.seed(5)
  out_t = distortion.preprocess_image(d, 'train', False, (96, 120), (30, 40), crop_size=(80, 100),
                                      image_distortion_fn=dict(random_brightness=True, random_saturation=True,
                                                               random_hue=True, random_contrast=True))
  assert out_t.shape == (3, 30, 40, 3) and out_t.dtype == torch.float32
  assert float(out_t.min()) >= 0.0 and float(out_t.max()) <= 1.0


def test_parallel_cheap_and_depth_distortions():
  """preprocessors/image_transformations.py:268-459: per-image parameter draws (Parallel), per-channel gamma (Cheap)
  and the depth-image distortions, against the oracle / numpy with the same drawn parameters."""
  from oracle import image_ops as oracle
  from tensor2robot_b200.preprocessors import image_ops
  from tensor2robot_b200.preprocessors import image_transformations as it
  rng = np.random.RandomState(5)
  x = rng.uniform(0.02, 0.98, (4, 24, 20, 3)).astype(np.float32)
  xg = torch.from_numpy(x).cuda()
  kwargs = dict(random_brightness=True, random_saturation=True, random_hue=True, random_contrast=True)
  it.seed(3)
  got = it.ApplyPhotometricImageDistortionsParallel(xg, **kwargs).cpu().numpy()
  it.seed(3)
  rec, _ = it.draw_photometric_params_parallel(4, **kwargs)
  assert len({float(v) for v in rec['brightness_delta']}) == 4               # one draw per image
  for i in range(4):
    want = oracle.distort(x[i:i + 1], brightness_delta=float(rec['brightness_delta'][i]),
                          saturation_scale=float(rec['saturation_scale'][i]), hue_delta=float(rec['hue_delta'][i]),
                          contrast_scale=float(rec['contrast_scale'][i]))
    assert np.abs(got[i:i + 1] - want).max() < 2e-5
  u8 = torch.from_numpy(rng.randint(0, 256, (4, 24, 20, 3)).astype(np.uint8)).cuda()
  it.seed(3)
  got_u8 = it.ApplyPhotometricImageDistortionsParallel(u8, **kwargs).cpu().numpy()
  want_u8 = np.concatenate([oracle.distort(u8[i:i + 1].cpu().numpy().astype(np.float32) / np.float32(255.0),
                                           brightness_delta=float(rec['brightness_delta'][i]),
                                           saturation_scale=float(rec['saturation_scale'][i]), hue_delta=float(rec['hue_delta'][i]),
                                           contrast_scale=float(rec['contrast_scale'][i])) for i in range(4)])
  assert np.abs(got_u8 - want_u8).max() < 2e-5
  with pytest.raises(NotImplementedError):
    it.ApplyPhotometricImageDistortionsParallel(xg, custom_distortion_fn=lambda im: im)

  it.seed(11)
  cheap = it.ApplyPhotometricImageDistortionsCheap(xg).cpu().numpy()
  it.seed(11)
  gammas = [it._RNG.uniform(0.5, 1.5) for _ in range(3)]                    # pylint: disable=protected-access
  assert np.abs(cheap - np.power(x.astype(np.float64), np.array(gammas))).max() < 2e-6

  depth = rng.uniform(0.1, 3.0, (2, 16, 12, 1)).astype(np.float32)
  dg = torch.from_numpy(depth).cuda()
  exact = image_ops.depth_distort(dg, 1.01, 0.0, 0.25, 2.5).cpu().numpy()
  np.testing.assert_allclose(exact, np.clip(np.float32(1.01) * depth, 0.25, 2.5), rtol=1e-6)
  big = torch.full((1, 256, 256, 1), 1.0, device='cuda')
  noisy = image_ops.depth_distort(big, 1.0, 0.05, 0.25, 2.5, seed=4).cpu().numpy().ravel()
  assert abs(noisy.mean() - 1.0) < 1e-3 and abs(noisy.std() - 0.05) < 1e-3
  assert not np.array_equal(noisy, image_ops.depth_distort(big, 1.0, 0.05, 0.25, 2.5, seed=5).cpu().numpy().ravel())
  it.seed(2)
  outs = it.ApplyDepthImageDistortions([dg] * 6, random_noise_level=0.05)
  changed = [not np.array_equal(o.cpu().numpy(), np.clip(depth, 0.25, 2.5)) for o in outs]
  assert any(changed) and not all(changed)                                   # the coin is drawn per list entry
  for o in outs:
    assert float(o.min()) >= 0.25 and float(o.max()) <= 2.5


@pytest.mark.gpu
def test_crop_convert_distort_vector_path_matches_scalar_path(tmp_path):
  """The 16-byte-vector kernel (crop width % 8 == 0) against the oracle for every source byte alignment
  (crop_x * 3 mod 16), and to one float ulp against the scalar kernel - including the Philox noise, which is
  indexed by pixel - run in a second process with T2R_DISABLE_VEC_CROP=1."""
  import os
  import subprocess
  import sys
  from oracle import image_ops as oracle
  from tensor2robot_b200.preprocessors import image_ops
  rng = np.random.RandomState(8)
  frames = rng.randint(0, 256, (3, 70, 150, 3)).astype(np.uint8)
  d = torch.from_numpy(frames).cuda()
  for crop_x in range(0, 17):                                    # all alignments of the first source byte
    p = image_ops.identity_params(3, 3, crop_x)
    ref = oracle.convert_image_dtype_f32(oracle.crop(frames, 3, crop_x, 63, 104))   # 63 rows: a partial last block
    out = image_ops.crop_convert_distort(d, (63, 104), p, torch.float32)
    assert np.array_equal(out.cpu().numpy(), ref), crop_x
    out_b = image_ops.crop_convert_distort(d, (63, 104), p, torch.bfloat16)
    assert torch.equal(out_b.cpu(), torch.from_numpy(ref).to(torch.bfloat16)), crop_x
  script = '''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from tensor2robot_b200.preprocessors import image_ops
frames = np.load(%r)
p = image_ops.identity_params(3, 2, 7)
p['noise_stddev'] = 0.05; p['brightness_delta'] = 0.03; p['saturation_scale'] = 1.2; p['hue_delta'] = -0.1; p['contrast_scale'] = 0.8
out = image_ops.crop_convert_distort(torch.from_numpy(frames).cuda(), (64, 136), p, torch.float32, seed=5, offset=3)
np.save(sys.argv[1], out.cpu().numpy())
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path / 'frames.npy'))
  np.save(str(tmp_path / 'frames.npy'), frames)
  outs = {}
  for name, env in (('vector', {}), ('scalar', {'T2R_DISABLE_VEC_CROP': '1'})):
    path = str(tmp_path / (name + '.npy'))
    subprocess.run([sys.executable, '-c', script, path], check=True, env=dict(os.environ, **env), timeout=300)
    outs[name] = np.load(path)
  # identical pixels, identical Philox draws; the compiler contracts the HSV arithmetic differently in the two kernels
  assert np.abs(outs['vector'] - outs['scalar']).max() < 5e-7
  assert outs['vector'].std() > 0.1
