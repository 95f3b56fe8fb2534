"""Schedules over the global step (utils/global_step_functions.py:26-123).  The reference builds graph tensors that read
the global-step variable; here a schedule is a callable `value = fn(global_step)`, the convention of
`models/optimizers.py` learning rates (evaluated on the host once per step and passed to the kernels as a scalar)."""
import numpy as np


def piecewise_linear(boundaries, values, name=None):
  """values[0] before boundaries[0], values[-1] from boundaries[-1] on, linear interpolation in between (:26-96)."""
  del name
  boundaries = np.asarray(boundaries, np.float64).reshape(-1)
  values = np.asarray(values, np.float64).reshape(-1)
  assert boundaries.size > 0, 'Need more than 0 boundaries'
  assert values.size > 0, 'Need more than 0 values'
  assert values.size == boundaries.size, 'boundaries and values must be of same size'

  def fn(global_step):
    x = float(global_step)
    # an unmet last boundary and an already met first one, carrying the end values (:64-75)
    b = np.concatenate([[min(x - 1, boundaries[0])], boundaries, [max(x + 1, boundaries[-1])]])
    v = np.concatenate([[values[0]], values, [values[-1]]])
    index = int(np.argmax(b > x)) if (b > x).any() else b.size      # first boundary not reached yet
    left, right = b[index - 1], b[index]
    a = (v[index] - v[index - 1]) / (right - left)
    return float(a * x + (v[index - 1] - a * left))

  return fn


def exponential_decay(initial_value=0.0001, decay_steps=10000, decay_rate=0.9, staircase=True):
  """tf.train.exponential_decay: initial_value * decay_rate ** (step / decay_steps), floored exponent if staircase."""

  def fn(global_step):
    p = float(global_step) / float(decay_steps)
    if staircase:
      p = np.floor(p)
    return float(initial_value * decay_rate ** p)

  return fn
