"""The per-replay-batch training step of the B200 engine (replaces the TF-Estimator train op).

Reference decomposition being replaced (models/abstract_model.py:694-755): validate_and_pack ->
inference_network_fn -> model_train_fn -> create_optimizer -> create_train_op, run by the Estimator's
Session.run loop (utils/train_eval.py:424-613).  Here one `CriticTrainStep.step` issues, on the
current CUDA stream and without any host synchronisation:

  crop + uint8->bf16 (+distortion)      1 HBM kernel         (t2r_models.py:297-308)
  critic forward                         tcgen05 convs ...    (networks.py:343-615)
  sigmoid + log loss (+ l2 term)         1 kernel             (t2r_models.py:229-239)
  backward                               dgrad/wgrad/BN kernels writing the flat gradient buffer
  gradient all-reduce (world_size > 1)   ONE NCCL all-reduce over the flat buffer
  optimizer + EMA + bf16 refresh         1 fused kernel       (optimizer_builder.py:25-96)
"""
import numpy as np
import torch
import torch.distributed as dist

from tensor2robot_b200 import nn
from tensor2robot_b200.preprocessors import image_ops


def reduce_gradients(flat_grad, world_size=None, group=None):
  """The ONE collective of a data-parallel step: sum the flat gradient buffer over the
  replicas (NCCL on GPU tensors, gloo on CPU tensors in the tests) and return the factor the fused
  optimizer kernel multiplies the gradient with (1 / world_size).  BatchNorm statistics stay
  per-replica, like the reference's towers."""
  if world_size is None:
    world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
  if world_size <= 1:
    return 1.0
  dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
  return 1.0 / world_size


class GradientReducer(object):
  """The data-parallel exchange of a training step (replaces the SyncReplicas / tower wrappers of
  models/abstract_model.py:864-870): the flat gradient buffer is all-reduced in a few buckets, each launched as
  soon as every variable in it has its gradient (nn.Variable.grad_ready), so the collective of the late layers
  overlaps the backward pass of the early ones instead of sitting behind the last weight gradient.

  Buckets are contiguous ranges of the flat buffer.  The buffer is laid out [regularised weights | the rest], each
  part in variable-creation (= forward) order, so a range near the end of a part is complete early in the backward
  pass: the weights are cut into `n_buckets` ranges of similar size, the small remainder (batch-norm parameters,
  biases, fp32 layers) is one more.  torch.distributed's NCCL backend runs each collective on its own stream after
  everything already queued on the launching stream; finish() makes the launching stream wait for all of them.
  Returns the 1 / world_size factor the fused optimizer kernel applies.  Sums are fp32 (a bf16 wire format would
  save < 0.1 ms of NVLink time per step at 96 MB and round every replica's gradient)."""

  def __init__(self, vs, world_size=None, group=None, n_buckets=3):
    if world_size is None:
      world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    self.vs, self.world_size, self.group = vs, world_size, group
    self.buckets = []          # [start, end, pending variable names]
    self._of = {}
    self._work = []
    self.launched_early = 0
    if world_size <= 1 or not vs.finalized:
      return
    train = [v for v in vs.vars.values() if v.trainable and getattr(v, 'offset', None) is not None]
    decay = sorted([v for v in train if v.regularize], key=lambda v: v.offset)
    rest = sorted([v for v in train if not v.regularize], key=lambda v: v.offset)
    target = max(1, sum(v.numel for v in decay) // max(1, n_buckets))
    groups, cur, size = [], [], 0
    for v in decay:
      cur.append(v)
      size += v.numel
      if size >= target and len(groups) < n_buckets - 1:
        groups.append(cur)
        cur, size = [], 0
    if cur:
      groups.append(cur)
    if rest:
      groups.append(rest)
    total = vs.flat_grad.numel()
    for i, g in enumerate(groups):
      start = g[0].offset
      nxt = groups[i + 1][0].offset if i + 1 < len(groups) else total
      self.buckets.append([start, nxt, set(v.name for v in g)])
      for v in g:
        self._of[v.name] = len(self.buckets) - 1

  def begin(self):
    """Call before backward(): arms the buckets and listens for finished gradients."""
    if self.world_size <= 1:
      return
    self._pending = [set(b[2]) for b in self.buckets]
    self._fired = [False] * len(self.buckets)
    self._work = []
    self.launched_early = 0
    self.vs.grad_listener = self._on_grad

  def _on_grad(self, var):
    i = self._of.get(var.name)
    if i is None or self._fired[i]:
      return
    self._pending[i].discard(var.name)
    if not self._pending[i]:
      self._fire(i)
      self.launched_early += 1

  def _fire(self, i):
    start, end, _ = self.buckets[i]
    self._fired[i] = True
    self._work.append(dist.all_reduce(self.vs.flat_grad[start:end], op=dist.ReduceOp.SUM, group=self.group,
                                      async_op=True))

  def finish(self):
    """Call after backward(): reduces whatever has not been launched, waits for everything."""
    if self.world_size <= 1:
      return 1.0
    self.vs.grad_listener = None
    for i in range(len(self.buckets)):
      if not self._fired[i]:
        self._fire(i)
    for work in self._work:
      work.wait()
    self._work = []
    return 1.0 / self.world_size


def optimization_step(vs, optimizer, global_step, loss=None, task_losses=None, reducer=None):
  """The tail every training step shares (the reference's train op, models/abstract_model.py:335-381):
  zero the flat gradient buffer, backward (or PCGrad's per-task backward passes), bucketed all-reduce overlapping
  the backward pass, ONE fused optimizer (+l2, +EMA, +bf16 refresh) kernel, refresh of the data-gradient weight
  packs.  Used by CriticTrainStep.step and AbstractT2RModel.train_step so the two cannot drift."""
  if reducer is not None:
    reducer.begin()
  if task_losses and hasattr(optimizer, 'compute_gradients'):
    # research/qtopt/pcgrad.py:99-121 (use_collection_losses): one backward per task loss, projected gradients;
    # the projection needs every task gradient, so the reducer only runs at the end
    if reducer is not None:
      vs.grad_listener = None
    optimizer.compute_gradients(list(task_losses), vs)
  else:
    vs.zero_grad()
    loss.backward()
  grad_scale = reducer.finish() if reducer is not None else 1.0
  optimizer.apply_gradients(vs, global_step, grad_scale)
  vs.sync_compute_copies(after_optimizer=True)


def shard_for_rank():
  """(rank, world_size) of this process for record sharding (files[rank::world])."""
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(), dist.get_world_size()
  return 0, 1


class CriticTrainStep(object):
  """Owns the variables, optimizer state and RNG of one data-parallel replica of a Q-critic."""

  def __init__(self, critic, optimizer, input_hw=(512, 640), target_hw=(472, 472), device='cuda', seed=0,
               world_size=1, rank=0, distort=None):
    self.critic = critic
    self.optimizer = optimizer
    self.optimizer.l2_regularization = critic.l2_regularization
    self.vs = nn.VariableStore(device, seed=seed)
    self.input_hw, self.target_hw = input_hw, target_hw
    self.world_size, self.rank = world_size, rank
    self.global_step = 0
    self.distort = distort or {}
    self._rng = np.random.RandomState(seed * 9973 + rank)
    self._built = False
    self.reducer = None

  # -- preprocessing (DefaultGrasping44ImagePreprocessor._preprocess_fn, t2r_models.py:277-308) ----
  def preprocess(self, images_u8, training=True, out_dtype=torch.bfloat16):
    n = images_u8.shape[0]
    (ih, iw), (th, tw) = images_u8.shape[1:3], self.target_hw
    if training:   # RandomCropImages: ONE offset pair per batch (image_transformations.py:25-59)
      oy, ox = int(self._rng.randint(0, ih - th + 1)), int(self._rng.randint(0, iw - tw + 1))
    else:          # CenterCropImages (image_transformations.py:62-101)
      oy, ox = (ih - th) // 2, (iw - tw) // 2
    params = image_ops.identity_params(n, oy, ox)
    seed = offset = 0
    if training and self.distort:
      d = self.distort
      if d.get('random_brightness'):
        params['brightness_delta'] = self._rng.uniform(-d.get('max_delta_brightness', 0.125),
                                                       d.get('max_delta_brightness', 0.125))
      if d.get('random_saturation'):
        params['saturation_scale'] = self._rng.uniform(d.get('lower_saturation', 0.5), d.get('upper_saturation', 1.5))
      if d.get('random_hue'):
        params['hue_delta'] = self._rng.uniform(-d.get('max_delta_hue', 0.2), d.get('max_delta_hue', 0.2))
      if d.get('random_contrast'):
        params['contrast_scale'] = self._rng.uniform(d.get('lower_contrast', 0.5), d.get('upper_contrast', 1.5))
      level = d.get('random_noise_level', 0.0)
      if level and not self._rng.uniform() > d.get('random_noise_apply_probability', 0.5):
        params['noise_stddev'] = level
        seed, offset = int(self._rng.randint(0, 2**31 - 1)), self.global_step
    return image_ops.crop_convert_distort(images_u8, (th, tw), params, out_dtype, seed, offset)

  # -- build -------------------------------------------------------------------------------------
  def build(self, images_u8, actions):
    """Creates the variables with a 2-sample inference pass, consolidates them into flat buffers and
    (world_size > 1) broadcasts rank 0's initial values."""
    if self._built:
      return
    with torch.no_grad(), nn.variable_store(self.vs):
      x = self.preprocess(images_u8[:2], training=False)
      self.critic.model((None, x), actions[:2], is_training=False)
    self.vs.finalize()
    if self.world_size > 1:
      dist.broadcast(self.vs.flat, src=0)
      dist.broadcast(self.vs.state_flat, src=0)
      self.vs.sync_compute_copies()
    self.reducer = GradientReducer(self.vs, self.world_size)
    self._built = True

  # -- one step ----------------------------------------------------------------------------------
  def step(self, images_u8, actions, reward):
    """images_u8 [B,H,W,3] uint8, actions [B,10] f32, reward [B,1] f32 - all CUDA.  Returns the
    device scalar loss = log_loss + l2 regularisation (no host sync)."""
    if not self._built:
      self.build(images_u8, actions)
    vs = self.vs
    with nn.variable_store(vs):
      x = self.preprocess(images_u8, training=True)
      logits, _ = self.critic.model((None, x), actions, is_training=True)
      loss, _ = nn.sigmoid_log_loss(logits, reward)
      total = loss.detach() + nn.l2_regularization_loss(self.critic.l2_regularization, vs)
      optimization_step(vs, self.optimizer, self.global_step, loss, reducer=self.reducer)
    self.global_step += 1
    return total

  @torch.no_grad()
  def predict(self, images_u8, actions, high_precision=False):
    """PREDICT mode: centre crop, moving-average BN, q_predicted [B] or [B, A].  high_precision: fp32 activations and
    bf16x3 convolutions (nn.high_precision), Q within 1e-3 relative of an fp32 evaluation."""
    with nn.variable_store(self.vs):
      if high_precision:
        with nn.high_precision():
          x = self.preprocess(images_u8, training=False, out_dtype=torch.float32)
          _, end_points = self.critic.model((None, x), actions, is_training=False)
      else:
        x = self.preprocess(images_u8, training=False)
        _, end_points = self.critic.model((None, x), actions, is_training=False)
    return end_points['predictions']


class LaggedTarget(object):
  """theta' of the Bellman target (SURVEY A-23): a lagged copy of the online critic's variables.

  QT-Opt evaluates max_a Q_theta'(s', a) with lagged target networks; the reference's nearest hook is the
  lagged checkpoint export of hooks/td3.py:37-132.  The copy lives in its own VariableStore (built by the
  same model code, hence the same flat layout): a refresh is two device-to-device copies of the flat
  buffers plus the bf16 recast.  `source='ema'` takes the MovingAverageOptimizer shadow instead of the raw
  weights (optimizer_builder.BuildOpt's use_avg_model_params)."""

  def __init__(self, step, update_every=100, source='online'):
    if source not in ('online', 'ema'):
      raise ValueError("source must be 'online' or 'ema'")
    self.step, self.update_every, self.source = step, int(update_every), source
    self.vs = nn.VariableStore(step.vs.device, seed=0)
    self.last_update = None

  def build(self, images_u8, actions):
    """Creates the target variables with the same 2-sample pass as CriticTrainStep.build."""
    step = self.step
    if not self.vs.finalized:
      with torch.no_grad(), nn.variable_store(self.vs):
        x = step.preprocess(images_u8[:2], training=False)
        step.critic.model((None, x), actions[:2], is_training=False)
      self.vs.finalize()
    self.update(force=True)

  def update(self, force=False):
    """Refreshes the copy when `update_every` optimizer steps have passed since the last refresh."""
    gs = int(self.step.global_step)
    if not force and self.last_update is not None and gs - self.last_update < self.update_every:
      return False
    ema = None
    if self.source == 'ema':
      ema = getattr(self.step.optimizer, '_ema', None)
      if ema is None:
        raise ValueError("source='ema' needs a MovingAverageOptimizer that has taken at least one step")
    self.vs.copy_values_from(self.step.vs, ema)
    self.last_update = gs
    return True


class CEMTargetComputer(object):
  """On-device cross-entropy-method maximisation of Q(s', a) and the Bellman target.

  Reference being replaced: policies/policies.py:133-184 runs utils/cross_entropy.py
  on the host and calls predictor.predict (a full graph run, image tower included, plus an H2D/D2H
  round trip) once per CEM iteration.  Here the state tower runs ONCE per transition, its output
  (`pool2` / block-layer-3 map) stays staged in HBM, and every iteration is: Philox sampling kernel
  -> one batched Q evaluation over [B*A] action samples against the staged features -> per-row
  elite refit kernel.  Semantics kept from the reference: initial mean 0 / stddev 1, ascending
  stable sort with the last `num_elites` kept, np.std(ddof=1), and the result is the arg-max over the
  LAST iteration's samples (policies.py:162).

  The Bellman target itself (y = r + gamma*(1-done)*max_a Q(s',a)) is not in the reference
  (SURVEY F3 / A-23): parity unpinned, validated by invariants.
  """

  def __init__(self, critic, vs, action_size=10, cem_samples=64, cem_iters=2, num_elites=10, seed=0,
               chunk=None, high_precision=False):
    """high_precision: evaluate the tower and every Q batch in nn.high_precision() (feed fp32 frames): the arg-max and
    the target then see Q within 1e-3 relative of an fp32 evaluation, at 3x the tensor work of the bf16 path."""
    self.critic, self.vs, self.chunk = critic, vs, chunk
    self.high_precision = high_precision
    self.action_size, self.cem_samples, self.cem_iters, self.num_elites = action_size, cem_samples, cem_iters, num_elites
    self.seed = seed
    self.calls = 0

  @torch.no_grad()
  def maximize(self, images_bf16):
    """images_bf16 [B,h,w,3] preprocessed next-state frames (fp32 with high_precision).  Returns (best_action [B,D],
    max_q [B], debug dict with the final mean/stddev)."""
    if self.high_precision and not nn.is_high_precision():
      with nn.high_precision():
        return self.maximize(images_bf16)
    import ctypes as C
    from tensor2robot_b200 import _lib
    dev = images_bf16.device
    b, a, d = images_bf16.shape[0], self.cem_samples, self.action_size
    st = _lib.current_stream_ptr()
    p = lambda t: C.c_void_p(t.data_ptr())
    mean = torch.zeros((b, d), dtype=torch.float32, device=dev)
    std = torch.ones((b, d), dtype=torch.float32, device=dev)
    samples = torch.empty((b, a, d), dtype=torch.float32, device=dev)
    best_q = torch.empty(b, dtype=torch.float32, device=dev)
    best_i = torch.empty(b, dtype=torch.int32, device=dev)
    q = torch.empty((b, a), dtype=torch.float32, device=dev)
    chunk = self.chunk or b
    with nn.variable_store(self.vs):
      with nn.variable_scope(self.critic.__class__.__name__):
        staged = self.critic.image_tower(images_bf16, False)      # the state tower runs ONCE
      for it in range(self.cem_iters):
        _lib.call('t2r_cem_sample', p(mean), p(std), p(samples), b, a, d, self.seed,
                  self.calls * self.cem_iters + it, st)
        # One batched Q evaluation per iteration; `chunk` bounds the [chunk*A, h, w, C] post-merge
        # activations (the ResNet-50 critic's merge map is 1.8 MB per action sample).
        for c0 in range(0, b, chunk):
          c1 = min(c0 + chunk, b)
          part = (staged[0][c0:c1], staged[1]) if isinstance(staged, tuple) else staged[c0:c1]
          _, ep = self.critic.model((None, None), samples[c0:c1], is_training=False, staged_features=part)
          q[c0:c1] = ep['predictions']
        _lib.call('t2r_cem_refit', p(samples), p(q), p(mean), p(std), p(best_q), p(best_i), b, a, d,
                  self.num_elites, st)
    self.calls += 1
    idx = best_i.long().view(b, 1, 1).expand(b, 1, d)
    best_action = samples.gather(1, idx).squeeze(1)
    return best_action, best_q, {'mean': mean, 'stddev': std, 'q': q, 'samples': samples}

  @torch.no_grad()
  def bellman_target(self, reward, done, max_q, gamma=0.9):
    import ctypes as C
    from tensor2robot_b200 import _lib
    target = torch.empty_like(max_q)
    _lib.call('t2r_bellman_target', C.c_void_p(reward.contiguous().data_ptr()),
              C.c_void_p(done.contiguous().data_ptr()), C.c_void_p(max_q.data_ptr()), float(gamma),
              C.c_void_p(target.data_ptr()), max_q.numel(), _lib.current_stream_ptr())
    return target


class BellmanCriticTrainStep(CriticTrainStep):
  """BASELINE config C3: the QT-Opt step with the CEM-maximised Bellman target computed inside it.

      y = r + gamma * (1 - done) * max_a Q_theta'(s', a)          (SURVEY A-23; arXiv 1806.10293, minimum viable form)

  theta' is a LaggedTarget of this step's critic; the arg-max is CEMTargetComputer (the next-state tower runs once,
  `cem_iters` x [B * cem_samples] batched post-merge passes against the staged features); y then replaces the
  `grasp_success` label of the supervised step (models/critic_model.py:171-192 consumes it as labels.reward).
  The reference has no such step (the target is computed "in a separate process (not open-sourced)",
  research/qtopt/README.md:9-12): parity of the target is unpinned by construction, its invariants are tested."""

  def __init__(self, critic, optimizer, gamma=0.9, cem_samples=64, cem_iters=2, num_elites=10, target_update_every=100,
               target_source='online', cem_chunk=None, high_precision_target=False, **kwargs):
    super(BellmanCriticTrainStep, self).__init__(critic, optimizer, **kwargs)
    self.gamma = gamma
    self.target = LaggedTarget(self, update_every=target_update_every, source=target_source)
    self.cem = CEMTargetComputer(critic, self.target.vs, action_size=10, cem_samples=cem_samples, cem_iters=cem_iters,
                                 num_elites=num_elites, seed=kwargs.get('rank', 0), chunk=cem_chunk,
                                 high_precision=high_precision_target)
    self.last_target = None

  def build(self, images_u8, actions):
    super(BellmanCriticTrainStep, self).build(images_u8, actions)
    if not self.target.vs.finalized:
      self.target.build(images_u8, actions)

  def step(self, images_u8, actions, reward, next_images_u8=None, done=None):
    """reward / done: [B, 1] or [B] float.  Without next-state frames this is the supervised step."""
    if next_images_u8 is None:
      return super(BellmanCriticTrainStep, self).step(images_u8, actions, reward)
    if not self._built:
      self.build(images_u8, actions)
    self.target.update()
    x_next = self.preprocess(next_images_u8, training=False,
                             out_dtype=torch.float32 if self.cem.high_precision else torch.bfloat16)
    _, max_q, _ = self.cem.maximize(x_next)
    y = self.cem.bellman_target(reward.reshape(-1).float(), done.reshape(-1).float(), max_q, self.gamma)
    self.last_target = y
    return super(BellmanCriticTrainStep, self).step(images_u8, actions, y.reshape(-1, 1))


def device_cem_selector(train_step, computer):
  """Robot-side use of the Bellman-target machinery: returns state -> (action [D] numpy, q float) for
  `policies.CEMPolicy(device_maximizer=...)`.  `state` is one uint8 camera frame [H, W, 3]; it is centre-cropped /
  converted by the train step's PREDICT preprocessing, the state tower runs once and CEM runs on the device
  (`CEMTargetComputer.maximize` at batch 1) - one H2D copy of the frame and one D2H copy of the action per call,
  instead of one predictor round trip per CEM iteration (policies/policies.py:164-182 in the reference)."""
  device = train_step.vs.device

  def select(state):
    frame = torch.from_numpy(np.ascontiguousarray(state)[None]).to(device, non_blocking=True)
    x = train_step.preprocess(frame, training=False)
    action, q, _ = computer.maximize(x)
    return action[0].float().cpu().numpy(), float(q[0])

  return select
