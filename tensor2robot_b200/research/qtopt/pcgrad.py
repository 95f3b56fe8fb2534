"""PCGrad - gradient surgery for multi-task learning (research/qtopt/pcgrad.py:29-244) - as an optimizer wrapper over
the engine's flat gradient buffers.  One backward pass per task loss fills a [tasks, parameters] buffer; the projection
itself is `t2r_pcgrad_project` (csrc/pcgrad.cu): Gram matrices of all variables in one pass, the sequential
projections on T x T coefficients, one combining pass - the per-variable implementation of the reference
(`use_per_variable_impl=True`, its default), with the same 1e-5 regulariser."""
import ctypes as C
import fnmatch
import random

import torch

from tensor2robot_b200 import _lib
from tensor2robot_b200.models import optimizers

MAX_TASKS = 8


def _p(t):
  return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class PCGrad(optimizers.Optimizer):
  """Wraps `optimizer_to_wrap`; `compute_gradients(losses, vs)` leaves the PCGrad gradient in vs.flat_grad."""

  def __init__(self, optimizer_to_wrap, allowlist=None, denylist=None, shuffle=True):
    super(PCGrad, self).__init__(None)
    self._optimizer = optimizer_to_wrap
    self._allowlist = allowlist
    self._denylist = denylist
    self._shuffle = shuffle        # the reference shuffles the task order on every graph construction
    self._tables = None

  @property
  def inner(self):
    return self._optimizer

  def uses_pcgrad(self, name):
    """pcgrad.py:81-97: allowlist (default everything) minus denylist, shell-style patterns on the variable name."""
    allow = ['*'] if self._allowlist is None else self._allowlist
    deny = [] if self._denylist is None else self._denylist
    return any(fnmatch.fnmatchcase(name, w) for w in allow) and not any(fnmatch.fnmatchcase(name, w) for w in deny)

  def _segment_tables(self, vs):
    if self._tables is None or self._tables[0] is not vs:
      train = vs.trainable_variables()
      off = torch.tensor([v.offset for v in train], dtype=torch.int64, device=vs.device)
      length = torch.tensor([v.numel for v in train], dtype=torch.int64, device=vs.device)
      use = torch.tensor([1 if self.uses_pcgrad(v.name) else 0 for v in train], dtype=torch.uint8, device=vs.device)
      self._tables = (vs, off, length, use, len(train))
    return self._tables[1:]

  def project(self, vs, task_grads):
    """task_grads: fp32 CUDA [T, vs.flat.numel()] per-task gradients -> vs.flat_grad (in place)."""
    if not task_grads.is_cuda:
      raise _lib.T2RError('PCGrad.project: gradients are on %s; the B200 engine has no CPU path' % task_grads.device)
    t, n = task_grads.shape
    if n != vs.flat.numel() or task_grads.dtype != torch.float32 or not task_grads.is_contiguous():
      raise ValueError('task_grads must be a contiguous fp32 [tasks, %d] tensor' % vs.flat.numel())
    if not 1 <= t <= MAX_TASKS:
      raise ValueError('PCGrad supports 1..%d task losses, got %d' % (MAX_TASKS, t))
    off, length, use, count = self._segment_tables(vs)
    gram = vs.scratch('pcgrad_gram', count * MAX_TASKS * MAX_TASKS, torch.float32)
    coef = vs.scratch('pcgrad_coef', count * MAX_TASKS, torch.float32)
    stream = _lib.current_stream_ptr()
    _lib.call('t2r_pcgrad_project', _p(task_grads), t, n, _p(off), _p(length), _p(use), count, 1e-5, _p(gram), _p(coef),
              _p(vs.flat_grad), stream)
    return vs.flat_grad

  def compute_gradients(self, losses, vs):
    """losses: list of scalar task losses that share one autograd graph (pcgrad.py:99-121)."""
    if not isinstance(losses, (list, tuple)):
      raise AssertionError('The loss is not a list: %s' % type(losses))
    losses = list(losses)
    if self._shuffle:
      random.shuffle(losses)
    buf = vs.scratch('pcgrad_task_grads', len(losses) * vs.flat.numel(), torch.float32)
    task_grads = buf[:len(losses) * vs.flat.numel()].view(len(losses), vs.flat.numel())
    for t, loss in enumerate(losses):
      vs.zero_grad()
      loss.backward(retain_graph=t + 1 < len(losses))
      task_grads[t].copy_(vs.flat_grad)
    return self.project(vs, task_grads)

  def apply_gradients(self, vs, global_step, grad_scale=1.0, ema=None, ema_decay=0.0):
    self._optimizer.l2_regularization = self.l2_regularization
    return self._optimizer.apply_gradients(vs, global_step, grad_scale, ema, ema_decay)

  def learning_rate(self, global_step):
    return self._optimizer.learning_rate(global_step)

  def slot_names(self):
    return self._optimizer.slot_names()

  def state_dict(self):
    return self._optimizer.state_dict()

  def load_state_dict(self, state, vs):
    return self._optimizer.load_state_dict(state, vs)
