"""Generates tests/golden/cem_golden.npz by running the REFERENCE's own utils/cross_entropy.py
(imported from /root/reference; it needs only numpy and six) on seeded inputs.

  python tests/golden/make_cem_golden.py

The reference's sample_fn is replaced by a recorded normal stream so that the run is reproducible
without numpy's global RNG state; everything else (sorting, elite selection, update, argmax) is the
reference's code.  Cases mirror CEMPolicy defaults (64 samples, 10 elites; policies.py:110-116).
"""
import importlib.util
import os
import sys

import numpy as np

REF = '/root/reference/utils/cross_entropy.py'
spec = importlib.util.spec_from_file_location('ref_cross_entropy', REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

out = {}
for case, (num_samples, num_elites, iters, dim, seed) in enumerate([(64, 10, 3, 10, 0), (64, 10, 2, 2, 1),
                                                                     (16, 4, 4, 3, 2)]):
  rng = np.random.RandomState(seed)
  noise = rng.standard_normal((iters, num_samples, dim))
  target = rng.uniform(-1, 1, dim)
  it = {'i': 0}
  log_samples, log_values, log_mean, log_std = [], [], [], []

  def sample_fn(mean, stddev):
    s = mean + stddev * noise[it['i']]
    it['i'] += 1
    return s

  def objective_fn(samples):
    v = -np.sum((np.asarray(samples) - target)**2, axis=1)
    v = np.round(v, 1)                      # coarse rounding creates ties: exercises sort stability
    log_samples.append(np.asarray(samples).copy())
    log_values.append(v.copy())
    return v

  def update_fn(params, elite_samples):
    del params
    p = {'mean': np.mean(elite_samples, axis=0), 'stddev': np.std(elite_samples, axis=0, ddof=1)}
    log_mean.append(p['mean'].copy())
    log_std.append(p['stddev'].copy())
    return p

  samples, values, final = ref.CrossEntropyMethod(
      sample_fn, objective_fn, update_fn, {'mean': np.zeros(dim), 'stddev': np.ones(dim)}, num_elites,
      num_iterations=iters)
  idx = int(np.argmax(values))
  p = 'case%d_' % case
  out[p + 'noise'] = noise
  out[p + 'target'] = target
  out[p + 'num_elites'] = num_elites
  out[p + 'samples'] = np.stack(log_samples)
  out[p + 'values'] = np.stack(log_values)
  out[p + 'mean'] = np.stack(log_mean)
  out[p + 'stddev'] = np.stack(log_std)
  out[p + 'best_index'] = idx
  out[p + 'best_action'] = np.asarray(samples)[idx]

# NormalCrossEntropyMethod end to end with numpy's global RNG (reference :110-154)
np.random.seed(123)
mean, std = ref.NormalCrossEntropyMethod(lambda s: -np.sum((s - 0.5)**2, axis=1), np.zeros(4), np.ones(4), 32, 6, 3)
out['normal_mean'], out['normal_stddev'] = mean, std

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'cem_golden.npz')
np.savez_compressed(dst, **out)
print('wrote', dst, os.path.getsize(dst), 'bytes')
