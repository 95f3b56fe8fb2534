"""Optimizers of the B200 engine, with the factory surface of the reference's models/optimizers.py.

Each optimizer is a thin host object around ONE fused kernel launch over the VariableStore's flat
buffers (tensor2robot_b200/csrc/loss_cem_optim.cu): gradient unscale (1/world_size) + slim l2
regulariser gradient + update + EMA (MovingAverageOptimizer) + bf16 weight refresh.

Reference: models/optimizers.py:26-58 (learning-rate fns), :61-129 (optimizer factories),
:132-159 (MovingAverageOptimizer / swapping saver); research/qtopt/optimizer_builder.py:25-96.
TF semantics restated: Adam uses the "epsilon hat" form
lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps); Momentum has no dampening/nesterov.
"""
import ctypes as C
import math

import torch

from tensor2robot_b200 import _lib


def _p(t):
  return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
  return _lib.current_stream_ptr()


# ---- learning-rate functions (called with the global step) --------------------------------------
def create_constant_learning_rate(initial_learning_rate=0.0001):
  """Returns the configured constant initial_learning_rate (models/optimizers.py:26-29)."""
  return initial_learning_rate


def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False):
  """tf.train.exponential_decay."""
  p = global_step / float(decay_steps)
  if staircase:
    p = math.floor(p)
  return learning_rate * decay_rate**p


def create_exp_decaying_learning_rate(initial_learning_rate=0.0001, decay_steps=10000, decay_rate=0.9,
                                      staircase=True):
  """A learning rate decaying exponentially with the global step (models/optimizers.py:32-58).
  Returns a callable of the global step (the reference returns a tensor bound to it)."""
  return lambda global_step: exponential_decay(initial_learning_rate, global_step, decay_steps, decay_rate,
                                               staircase)


def _lr_value(learning_rate, global_step):
  return float(learning_rate(global_step)) if callable(learning_rate) else float(learning_rate)


# ---- optimizers ---------------------------------------------------------------------------------
class Optimizer(object):
  """Base: owns fp32 slot buffers shaped like the store's flat parameter buffer."""

  def __init__(self, learning_rate):
    self._learning_rate = learning_rate
    self._slots = {}
    self.l2_regularization = 0.0   # slim l2_regularizer scale for the store's regularised prefix

  def _slot(self, vs, name):
    t = self._slots.get(name)
    if t is None or t.numel() != vs.flat.numel():
      t = torch.zeros_like(vs.flat)
      self._slots[name] = t
    return t

  def slot_names(self):
    return sorted(self._slots)

  def state_dict(self):
    return {k: v.detach().cpu() for k, v in self._slots.items()}

  def load_state_dict(self, state, vs):
    for k, v in state.items():
      self._slot(vs, k).copy_(v.to(vs.device))

  def learning_rate(self, global_step):
    return _lr_value(self._learning_rate, global_step)

  def apply_gradients(self, vs, global_step, grad_scale=1.0, ema=None, ema_decay=0.0):
    raise NotImplementedError


class GradientDescentOptimizer(Optimizer):

  def apply_gradients(self, vs, global_step, grad_scale=1.0, ema=None, ema_decay=0.0):
    accum = self._slot(vs, 'accum')   # momentum 0 => accum == grad, w -= lr*grad
    _lib.call('t2r_momentum_step', _p(vs.flat), _p(vs.flat_grad), _p(accum), _p(ema), _p(vs.flat_bf16),
              vs.flat.numel(), vs.n_decay, self.learning_rate(global_step), 0.0, self.l2_regularization,
              grad_scale, ema_decay, _stream())


class MomentumOptimizer(Optimizer):
  """tf.train.MomentumOptimizer (use_nesterov=False)."""

  def __init__(self, learning_rate, momentum=0.9):
    super(MomentumOptimizer, self).__init__(learning_rate)
    self._momentum = momentum

  def apply_gradients(self, vs, global_step, grad_scale=1.0, ema=None, ema_decay=0.0):
    accum = self._slot(vs, 'momentum')
    _lib.call('t2r_momentum_step', _p(vs.flat), _p(vs.flat_grad), _p(accum), _p(ema), _p(vs.flat_bf16),
              vs.flat.numel(), vs.n_decay, self.learning_rate(global_step), self._momentum,
              self.l2_regularization, grad_scale, ema_decay, _stream())


class AdamOptimizer(Optimizer):
  """tf.train.AdamOptimizer."""

  def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
    super(AdamOptimizer, self).__init__(learning_rate)
    self._beta1, self._beta2, self._epsilon = beta1, beta2, epsilon

  def apply_gradients(self, vs, global_step, grad_scale=1.0, ema=None, ema_decay=0.0):
    m, v = self._slot(vs, 'm'), self._slot(vs, 'v')
    _lib.call('t2r_adam_step', _p(vs.flat), _p(vs.flat_grad), _p(m), _p(v), _p(ema), _p(vs.flat_bf16),
              vs.flat.numel(), vs.n_decay, self.learning_rate(global_step), self._beta1, self._beta2,
              self._epsilon, int(global_step) + 1, self.l2_regularization, grad_scale, ema_decay, _stream())


class RMSPropOptimizer(Optimizer):
  """tf.train.RMSPropOptimizer (centered=False): the `rms` slot starts at one, `momentum` at zero."""

  def __init__(self, learning_rate, decay=0.9, momentum=0.0, epsilon=1e-10):
    super(RMSPropOptimizer, self).__init__(learning_rate)
    self._decay, self._momentum, self._epsilon = decay, momentum, epsilon

  def apply_gradients(self, vs, global_step, grad_scale=1.0, ema=None, ema_decay=0.0):
    fresh = 'rms' not in self._slots or self._slots['rms'].numel() != vs.flat.numel()
    rms, mom = self._slot(vs, 'rms'), self._slot(vs, 'momentum')
    if fresh:
      rms.fill_(1.0)
    _lib.call('t2r_rmsprop_step', _p(vs.flat), _p(vs.flat_grad), _p(rms), _p(mom), _p(ema), _p(vs.flat_bf16),
              vs.flat.numel(), vs.n_decay, self.learning_rate(global_step), self._decay, self._momentum, self._epsilon,
              self.l2_regularization, grad_scale, ema_decay, _stream())


class MovingAverageOptimizer(Optimizer):
  """contrib.opt.MovingAverageOptimizer: wraps an optimizer and keeps an EMA shadow of every
  trainable variable (the shadow is what the reference's swapping saver writes to checkpoints)."""

  def __init__(self, opt, average_decay=0.9999):
    super(MovingAverageOptimizer, self).__init__(None)
    self._opt = opt
    self._average_decay = average_decay
    self._ema = None

  @property
  def inner(self):
    return self._opt

  @property
  def l2_regularization(self):
    return self._opt.l2_regularization

  @l2_regularization.setter
  def l2_regularization(self, v):
    if hasattr(self, '_opt'):
      self._opt.l2_regularization = v

  def learning_rate(self, global_step):
    return self._opt.learning_rate(global_step)

  def shadow(self, vs):
    if self._ema is None or self._ema.numel() != vs.flat.numel():
      self._ema = vs.flat.clone()   # tf ExponentialMovingAverage initialises the shadow to the value
    return self._ema

  def apply_gradients(self, vs, global_step, grad_scale=1.0, ema=None, ema_decay=0.0):
    self._opt.apply_gradients(vs, global_step, grad_scale, self.shadow(vs), self._average_decay)

  def state_dict(self):
    d = {'opt/' + k: v for k, v in self._opt.state_dict().items()}
    if self._ema is not None:
      d['ema'] = self._ema.detach().cpu()
    return d

  def load_state_dict(self, state, vs):
    self._opt.load_state_dict({k[4:]: v for k, v in state.items() if k.startswith('opt/')}, vs)
    if 'ema' in state:
      self.shadow(vs).copy_(state['ema'].to(vs.device))


def moving_average_shadow(optimizer):
  """The EMA shadow buffer (flat fp32, the layout of VariableStore.flat) of a MovingAverageOptimizer, looking
  through wrappers that hold their optimizer as `_optimizer` / `inner` (PCGrad); None when there is none yet."""
  seen = 0
  while optimizer is not None and seen < 4:
    if isinstance(optimizer, MovingAverageOptimizer):
      return optimizer._ema   # pylint: disable=protected-access
    optimizer = getattr(optimizer, '_optimizer', None)
    seen += 1
  return None


# ---- factories with the reference's names --------------------------------------------------------
def default_create_optimizer_fn(use_summaries, learning_rate=1e-4):
  del use_summaries
  return AdamOptimizer(learning_rate)


def create_adam_optimizer(learning_rate_fn=create_constant_learning_rate):
  def create_optimizer_fn(use_summaries):
    del use_summaries
    return AdamOptimizer(learning_rate=learning_rate_fn())
  return create_optimizer_fn


def create_gradient_descent_optimizer(learning_rate_fn=create_constant_learning_rate):
  def create_optimizer_fn(use_summaries):
    del use_summaries
    return GradientDescentOptimizer(learning_rate=learning_rate_fn())
  return create_optimizer_fn


def create_momentum_optimizer(learning_rate_fn=create_constant_learning_rate, momentum=0.9):
  def create_optimizer_fn(use_summaries):
    del use_summaries
    return MomentumOptimizer(learning_rate=learning_rate_fn(), momentum=momentum)
  return create_optimizer_fn


def create_moving_average_optimizer(optimizer, average_decay=0.999):
  return MovingAverageOptimizer(optimizer, average_decay=average_decay)
