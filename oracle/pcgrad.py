"""TEST INFRASTRUCTURE ONLY - numpy restatement of research/qtopt/pcgrad.py:123-154
(`_compute_projected_grads_per_variable`): per variable, every task gradient is projected sequentially against all
task gradients (itself included) and the projected gradients are summed.  Pinned by the known-answer vectors of
research/qtopt/pcgrad_test.py:35-129 (tests/test_pcgrad.py)."""
import fnmatch

import numpy as np


def project_variable(task_grads, eps=1e-5):
  """task_grads: list of T arrays (one variable's gradient per task) -> the PCGrad gradient of the variable."""
  task_grads = [np.asarray(g, np.float64) for g in task_grads]
  var_grad = 0
  for task_grad in task_grads:
    grad = task_grad
    for inner in task_grads:
      proj_direction = np.sum(grad * inner) / (np.sum(inner * inner) + eps)
      grad = grad - min(proj_direction, 0.) * inner
    var_grad = var_grad + grad
  return var_grad


def uses_pcgrad(name, allowlist=None, denylist=None):
  """pcgrad.py:81-97."""
  allow = ['*'] if allowlist is None else allowlist
  deny = [] if denylist is None else denylist
  return any(fnmatch.fnmatchcase(name, w) for w in allow) and not any(fnmatch.fnmatchcase(name, w) for w in deny)


def compute_gradients(task_var_grads, names, allowlist=None, denylist=None):
  """task_var_grads[t][name] -> {name: gradient}; variables outside the lists get the gradient of the summed loss."""
  out = {}
  for n in names:
    gs = [tg[n] for tg in task_var_grads]
    out[n] = project_variable(gs) if uses_pcgrad(n, allowlist, denylist) else np.sum(np.asarray(gs, np.float64), 0)
  return out
