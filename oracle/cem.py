"""ORACLE (test infrastructure, never imported by the product): numpy restatement of the reference's
cross-entropy method and of the engine's Philox sampler.

  cross_entropy_method      utils/cross_entropy.py:30-107  (sort ascending, keep last num_elites)
  normal update_fn          utils/cross_entropy.py:137-142 / policies/policies.py:145-150
                            (np.mean, np.std(ddof=1) over the elite samples)
  cem_policy_argmax         policies/policies.py:133-169   (mu=0, sigma=1, argmax over LAST samples)

PINNED: tests/golden/cem_golden.npz holds inputs/outputs produced by the reference's own
utils/cross_entropy.py imported from /root/reference (tests/golden/make_cem_golden.py); this
restatement is checked against it in tests/test_cem.py.
"""
import operator

import numpy as np


def cross_entropy_method(sample_fn, objective_fn, update_fn, initial_params, num_elites, num_iterations=1,
                         threshold_to_terminate=None):
  """List-valued sample batches only (the form CEMPolicy uses)."""
  params = initial_params
  samples = values = None
  for _ in range(num_iterations):
    samples = sample_fn(**params)
    values = objective_fn(samples)
    order = [i for i, _ in sorted(enumerate(values), key=operator.itemgetter(1))]   # stable ascending
    elites = [samples[i] for i in order][-num_elites:]
    params = update_fn(params, elites)
    if threshold_to_terminate is not None and max(values) > threshold_to_terminate:
      break
  return samples, values, params


def normal_update_fn(params, elite_samples):
  del params
  return {'mean': np.mean(elite_samples, axis=0), 'stddev': np.std(elite_samples, axis=0, ddof=1)}


def refit_rows(samples, values, num_elites):
  """Batched form of one CEM update: samples [B,A,D], values [B,A] -> mean, std [B,D], best value,
  first arg-max index (np.argmax) per row."""
  b, a, d = samples.shape
  mean = np.zeros((b, d), np.float64)
  std = np.zeros((b, d), np.float64)
  for i in range(b):
    order = np.argsort(values[i], kind='stable')
    elites = samples[i][order[-num_elites:]].astype(np.float64)
    mean[i] = elites.mean(0)
    std[i] = elites.std(0, ddof=1)
  return mean, std, values.max(1), values.argmax(1)


# ---- Philox4x32-10 + Box-Muller exactly as tensor2robot_b200/csrc/philox.cuh -------------------
_M0, _M1, _W0, _W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(seed, index, offset):
  """index: uint64 array.  Returns uint32 array [..., 4]."""
  index = np.asarray(index, np.uint64)
  k0 = np.uint32(seed & 0xFFFFFFFF)
  k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
  c0 = (index & np.uint64(0xFFFFFFFF)).astype(np.uint32)
  c1 = (index >> np.uint64(32)).astype(np.uint32)
  c2 = np.full_like(c0, np.uint32(offset & 0xFFFFFFFF))
  c3 = np.full_like(c0, np.uint32((offset >> 32) & 0xFFFFFFFF))
  with np.errstate(over='ignore'):
    for _ in range(10):
      p0 = _M0 * c0.astype(np.uint64)
      p1 = _M1 * c2.astype(np.uint64)
      n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
      n1 = p1.astype(np.uint32)
      n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
      n3 = p0.astype(np.uint32)
      c0, c1, c2, c3 = n0, n1, n2, n3
      k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
      k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
  return np.stack([c0, c1, c2, c3], -1)


def _u01(x):
  return ((x >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)


def philox_normal(seed, offset, count):
  """`count` standard normals in the order the cem_sample kernel emits them (4 per counter)."""
  quads = (count + 3) // 4
  r = philox4x32_10(seed, np.arange(quads, dtype=np.uint64), offset)
  out = np.zeros((quads, 4), np.float32)
  for pair in range(2):
    u1, u2 = _u01(r[:, 2 * pair]), _u01(r[:, 2 * pair + 1])
    rad = np.sqrt(np.float32(-2.0) * np.log(u1))
    ang = np.float32(6.28318530717958647692) * u2
    out[:, 2 * pair] = rad * np.cos(ang)
    out[:, 2 * pair + 1] = rad * np.sin(ang)
  return out.reshape(-1)[:count]


def cem_sample(mean, stddev, num_samples, seed, offset):
  """samples[b,a,d] = mean[b,d] + stddev[b,d] * z  (t2r_cem_sample)."""
  b, d = mean.shape
  z = philox_normal(seed, offset, b * num_samples * d).reshape(b, num_samples, d)
  return (mean[:, None, :] + stddev[:, None, :] * z).astype(np.float32)
