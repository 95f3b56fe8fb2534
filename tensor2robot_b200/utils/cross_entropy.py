"""Cross-entropy method with the reference's host API (utils/cross_entropy.py:30-154).

This module is host control flow over arbitrary Python callables (sample_fn / objective_fn /
update_fn), exactly the contract policies.CEMPolicy drives; it is what a robot-side policy calls at
1-10 Hz.  The per-replay-batch hot path does NOT use it: training-time action maximisation runs on
the GPU in tensor2robot_b200.engine.CEMTargetComputer (Philox sampling, batched Q evaluation against
staged image features, elite refit - one kernel each).
"""
import numpy as np


def _take(batch, indices):
  return [batch[i] for i in indices]


def CrossEntropyMethod(sample_fn, objective_fn, update_fn, initial_params, num_elites, num_iterations=1,  # pylint: disable=invalid-name
                       threshold_to_terminate=None):
  """Maximises objective_fn with CEM.  Sample batches are lists of samples or dicts of such lists.

  Returns (final_samples, final_values, final_params): the samples and values of the LAST
  iteration and the parameters updated from its elites.
  """
  params = initial_params
  samples, values = None, None
  for _ in range(num_iterations):
    samples = sample_fn(**params)
    values = objective_fn(samples)
    # ascending, stable: among equal values the later sample ranks higher
    order = sorted(range(len(values)), key=lambda i: values[i])
    elite_idx = order[-num_elites:]
    if isinstance(samples, dict):
      elites = {k: _take(v, elite_idx) for k, v in samples.items()}
    else:
      elites = _take(samples, elite_idx)
    params = update_fn(params, elites)
    if threshold_to_terminate is not None and max(values) > threshold_to_terminate:
      break
  return samples, values, params


def NormalCrossEntropyMethod(objective_fn, mean, stddev, num_samples, num_elites, num_iterations=1):  # pylint: disable=invalid-name
  """CEM with a diagonal normal sampling distribution; returns the final (mean, stddev)."""
  size = np.broadcast(mean, stddev).size

  def sample_fn(mean, stddev):
    return mean + stddev * np.random.randn(num_samples, size)

  def update_fn(params, elite_samples):
    del params
    return {'mean': np.mean(elite_samples, axis=0), 'stddev': np.std(elite_samples, axis=0, ddof=1)}

  _, _, final = CrossEntropyMethod(sample_fn, objective_fn, update_fn, {'mean': mean, 'stddev': stddev},
                                   num_elites, num_iterations=num_iterations)
  return final['mean'], final['stddev']
