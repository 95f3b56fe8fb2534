"""train_eval_model: the training / evaluation driver of the B200 engine.

Keeps the entry point of the reference (utils/train_eval.py:424-613,
`train_eval_model(t2r_model, input_generator_train, input_generator_eval, max_train_steps, model_dir,
eval_steps, ...)`) and replaces tf.estimator.Estimator with a CUDA-stream driver:

  host batch (numpy) -> pinned staging buffers -> async H2D on a copy stream (double buffered,
  overlapping the previous step's compute) -> model.preprocessor.preprocess on the GPU ->
  model.train_step (forward, backward, NCCL all-reduce, fused optimizer) -> checkpoint cadence.

Checkpoints are torch files of {reference variable name: array in TF layout} + optimizer slots +
global_step, written as model_dir/model.ckpt-<step>.pt; training resumes from the newest.
"""
import glob
import logging
import os
import re
import sys

import numpy as np

import torch

from tensor2robot_b200.models import model_interface
from tensor2robot_b200.utils import tensorspec_utils

ModeKeys = model_interface.ModeKeys


def print_spec(tensor_spec):
  for key, value in tensorspec_utils.flatten_spec_structure(tensor_spec).items():
    logging.info('%s: %s', key, value)


def print_specification(t2r_model):
  """Logs the model / preprocessor specifications for all modes (:73-94)."""
  for mode in (ModeKeys.TRAIN, ModeKeys.EVAL, ModeKeys.PREDICT):
    logging.info('Preprocessor in feature specification (%s):', mode)
    print_spec(t2r_model.preprocessor.get_in_feature_specification(mode))
    logging.info('Model feature specification (%s):', mode)
    print_spec(t2r_model.get_feature_specification(mode))


def provide_input_generator_with_model_information(input_generator_or_generators, t2r_model, mode):
  """Sets the specifications and preprocess function on the input generator(s) (:97-129)."""
  if isinstance(input_generator_or_generators, (list, tuple)):
    for g in input_generator_or_generators:
      g.set_specification_from_model(t2r_model, mode)
  else:
    input_generator_or_generators.set_specification_from_model(t2r_model, mode)
  return input_generator_or_generators


def tfdata_image_decoder():
  from tensor2robot_b200.utils import tfdata
  return tfdata.image_decoder()


class DeviceStager(object):
  """Pinned host staging + async H2D on a dedicated copy stream, `depth` pinned slots.

  Hazards handled here (both would corrupt inputs silently):
    * a pinned slot is refilled only after the H2D copies that read it have completed (one event per slot,
      synchronised before the host memcpy into the slot);
    * the device tensors are allocated on the copy stream but consumed on `consumer_stream`: record_stream tells the
      caching allocator not to hand their memory to a later copy while the consumer's kernels may still read it.
  stage() may run on a producer thread (train_eval._batches does): it only touches the copy stream.  Large sources that
  are already page-locked are not copied on the host at all; their owner must not rewrite them before `depth` further
  batches have been staged."""

  def __init__(self, device, depth=3):
    self.device = torch.device(device)
    self.stream = torch.cuda.Stream(device=self.device)
    self.depth = depth
    self._pinned = [dict() for _ in range(depth)]
    self._copied = [None] * depth
    self._slot = 0
    self._pool = None

  _PARALLEL_COPY_BYTES = 32 << 20
  _COPY_THREADS = 8

  def _fill(self, buf, t):
    """Host memcpy into the pinned slot.  One core moves ~10 GB/s: a 0.5 GB frame batch would cost 50 ms, more than
    a BC-Z step, so large tensors are copied in byte ranges by a few threads with plain memmove (ctypes releases the
    GIL; torch's copy_ from several Python threads serialises on its intra-op thread pool: measured 12 GB/s)."""
    nbytes = t.numel() * t.element_size()
    if nbytes < self._PARALLEL_COPY_BYTES or not (t.is_contiguous() and buf.is_contiguous()):
      buf.copy_(t)
      return
    if self._pool is None:
      import concurrent.futures
      self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=self._COPY_THREADS)
    import ctypes
    dst, src = buf.data_ptr(), t.data_ptr()
    step = -(-nbytes // self._COPY_THREADS)
    step = (step + 4095) & ~4095
    ranges = [(o, min(step, nbytes - o)) for o in range(0, nbytes, step)]
    list(self._pool.map(lambda r: ctypes.memmove(dst + r[0], src + r[0], r[1]), ranges))

  def stage(self, struct, consumer_stream=None):
    """struct: flat {path: numpy | torch CPU tensor}.  Returns (device struct, ready event)."""
    index = self._slot
    slot = self._pinned[index]
    self._slot = (self._slot + 1) % self.depth
    if self._copied[index] is not None:
      self._copied[index].synchronize()
    if consumer_stream is None:
      consumer_stream = torch.cuda.current_stream(self.device)
    out = tensorspec_utils.TensorSpecStruct()
    with torch.cuda.stream(self.stream):
      for key, value in struct.items():
        if isinstance(value, torch.Tensor) and value.is_cuda:
          # already on the device: frames the split JPEG decoder produced on this (copy) stream
          value.record_stream(consumer_stream)
          out[key] = value
          continue
        t = value if isinstance(value, torch.Tensor) else torch.from_numpy(value)
        if t.dtype == torch.float64:
          t = t.float()
        if t.numel() * t.element_size() >= self._PARALLEL_COPY_BYTES and t.is_contiguous() and t.is_pinned():
          # the producer already parsed / decoded into page-locked memory (a record reader's output buffers, a
          # generator that owns pinned arrays): DMA straight from it; the slot keeps it alive until the copy is done
          slot[key] = t
          dev = t.to(self.device, non_blocking=True)
          dev.record_stream(consumer_stream)
          out[key] = dev
          continue
        buf = slot.get(key)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype or buf is t:
          buf = slot[key] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
        self._fill(buf, t)
        dev = buf.to(self.device, non_blocking=True)
        dev.record_stream(consumer_stream)
        out[key] = dev
      ready = torch.cuda.Event()
      ready.record(self.stream)
    self._copied[index] = ready
    return out, ready


def latest_checkpoint(model_dir):
  best, best_step = None, -1
  for path in glob.glob(os.path.join(model_dir, 'model.ckpt-*.pt')):
    m = re.search(r'model\.ckpt-(\d+)\.pt$', path)
    if m and int(m.group(1)) > best_step:
      best, best_step = path, int(m.group(1))
  return best


def _is_chief():
  return not (torch.distributed.is_available() and torch.distributed.is_initialized()) or \
      torch.distributed.get_rank() == 0


def _replace_into_place(tmp, path):
  """A checkpoint becomes visible under its final name only when complete: readers (other ranks, a restart after a
  crash, predictors polling model_dir) never see a truncated file."""
  os.replace(tmp, path)


def save_checkpoint(t2r_model, model_dir, keep_checkpoint_max=5):
  path = os.path.join(model_dir, 'model.ckpt-%d.pt' % t2r_model.global_step)
  if not _is_chief():
    return path          # replicas hold identical parameters: the chief writes (the Estimator's chief-only saver)
  os.makedirs(model_dir, exist_ok=True)
  tmp = path + '.tmp-%d' % os.getpid()
  torch.save(t2r_model.state_dict(), tmp)
  _replace_into_place(tmp, path)
  existing = sorted(glob.glob(os.path.join(model_dir, 'model.ckpt-*.pt')),
                    key=lambda p: int(re.search(r'-(\d+)\.pt$', p).group(1)))
  for old in existing[:-keep_checkpoint_max]:
    os.remove(old)
  return path


def save_tf_checkpoint(t2r_model, model_dir):
  """Writes the model variables (reference names, TF layouts) + global_step as a TensorFlow tensor-bundle checkpoint
  `model_dir/model.ckpt-<step>.{index,data-00000-of-00001}` with the `checkpoint` state file tf.train.latest_checkpoint
  reads: weights trained here load into the reference (tf.train.load_checkpoint / init_from_checkpoint) by name."""
  from tensor2robot_b200.utils import tf_checkpoint
  name = 'model.ckpt-%d' % t2r_model.global_step
  prefix = os.path.join(model_dir, name)
  if not _is_chief():
    return prefix
  os.makedirs(model_dir, exist_ok=True)
  # with a MovingAverageOptimizer the variable names carry the AVERAGED values (the reference's swapping saver,
  # models/abstract_model.py:855-863): this is what exports, predictors and the lagged target load
  export = getattr(t2r_model, 'export_variables', None) or t2r_model.variable_store.export_tf
  tensors = dict(export())
  tensors['global_step'] = np.asarray(t2r_model.global_step, np.int64)
  tmp_prefix = os.path.join(model_dir, '.tmp-%d-%s' % (os.getpid(), name))
  tf_checkpoint.write_checkpoint(tmp_prefix, tensors)
  for suffix in ('.data-00000-of-00001', '.index'):      # the index last: it is what readers look for
    _replace_into_place(tmp_prefix + suffix, prefix + suffix)
  state = os.path.join(model_dir, 'checkpoint')
  with open(state + '.tmp', 'w') as f:
    f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (name, name))
  _replace_into_place(state + '.tmp', state)
  return prefix


class Prefetcher(object):
  """Runs a host-batch iterator on a background thread, `depth` batches ahead (tf.data's prefetch,
  utils/tfdata.py:629-689).  Record reading, tf.Example parsing and image decoding spend their time in
  C / C++ with the GIL released, so they overlap the main thread's kernel launches and the GPU step.
  Exceptions of the producer are re-raised in the consumer; close() (or exhausting / deleting the
  iterator) stops the thread."""
  _END = object()

  def __init__(self, iterable, depth=2):
    import queue
    import threading
    self._queue = queue.Queue(maxsize=max(1, depth))
    self._stop = threading.Event()
    self._thread = threading.Thread(target=self._run, args=(iterable,), daemon=True)
    self._thread.start()

  def _run(self, iterable):
    import queue
    try:
      for item in iterable:
        while not self._stop.is_set():
          try:
            self._queue.put(item, timeout=0.1)
            break
          except queue.Full:
            continue
        if self._stop.is_set():
          return
      item = self._END
    except BaseException as e:  # pylint: disable=broad-except
      item = e
    while not self._stop.is_set():
      try:
        self._queue.put(item, timeout=0.1)
        return
      except queue.Full:
        continue

  def __iter__(self):
    return self

  def __next__(self):
    item = self._queue.get()
    if item is self._END:
      self._queue.put(item)       # stay exhausted
      raise StopIteration
    if isinstance(item, BaseException):
      self._queue.put(item)
      raise item
    return item

  def close(self):
    self._stop.set()

  def __del__(self):
    self.close()


def _batches(input_generator, t2r_model, mode, device, prefetch=2):
  """Host batches (prefetched on a background thread) -> staged device batches -> preprocessed
  (features, labels)."""
  preprocessor = t2r_model.preprocessor
  consumer_stream = torch.cuda.current_stream(device)
  on_thread = bool(prefetch)
  # every queued batch holds a pinned slot, plus the one being filled and the one the GPU may still be copying
  stager = DeviceStager(device, depth=(prefetch + 2) if on_thread else 2)

  def host_batches():
    """The input generator's batches.  With the split JPEG decoder the generator itself launches kernels (IDCT / colour
    on the GPU): they go to the stager's copy stream, like the H2D copies, so that the whole record path - read, parse,
    Huffman threads, device half of the decode - runs ahead of the training step on the producer thread."""
    it = iter(input_generator.create_dataset(mode))
    while True:
      with torch.cuda.stream(stager.stream):
        try:
          item = next(it)
        except StopIteration:
          return
      yield item

  def staged_batches():
    """host batch -> pinned slot -> async H2D.  The 0.5 GB host memcpy of a 512-frame batch stays off the thread
    that launches kernels when this generator runs inside the Prefetcher."""
    if on_thread:
      torch.cuda.set_device(device)
    for features, labels in host_batches():
      merged = tensorspec_utils.TensorSpecStruct(
          [('f/' + k, v) for k, v in tensorspec_utils.flatten_spec_structure(features).items()])
      if labels is not None:
        for k, v in tensorspec_utils.flatten_spec_structure(labels).items():
          merged['l/' + k] = v
      yield stager.stage(merged, consumer_stream)

  source = Prefetcher(staged_batches(), prefetch) if on_thread else staged_batches()
  for staged, ready in source:
    consumer_stream.wait_event(ready)
    dev_features = tensorspec_utils.TensorSpecStruct([(k[2:], v) for k, v in staged.items() if k.startswith('f/')])
    dev_labels = tensorspec_utils.TensorSpecStruct([(k[2:], v) for k, v in staged.items() if k.startswith('l/')])
    yield preprocessor.preprocess(dev_features, dev_labels if len(dev_labels) else None, mode)


def train_eval_model(t2r_model=None, input_generator_train=None, input_generator_eval=None, max_train_steps=1000,
                     model_dir='/tmp/t2r_b200', eval_steps=100, eval_throttle_secs=600, create_exporters_fn=None,
                     export_generator=None, use_continuous_eval=True, train_hook_builders=None,
                     chief_train_hook_builders=None, eval_hook_builders=None, device=None, log_every_n_steps=100):
  """Trains (and evaluates) a T2R model.  Returns {'global_step', 'loss', 'eval'}.

  train_hook_builders / chief_train_hook_builders: tensor2robot_b200.hooks.HookBuilder objects (begin / before_step /
  after_step / end hooks).  Exporters and eval hooks of the Estimator world have no analogue here and must be None."""
  del eval_throttle_secs, use_continuous_eval
  if any(x is not None for x in (create_exporters_fn, export_generator, eval_hook_builders)):
    raise NotImplementedError('SavedModel exporters / eval hooks are outside the B200 engine (SURVEY 2)')
  hooks = []
  is_chief = _is_chief()
  distributed = torch.distributed.is_available() and torch.distributed.is_initialized()
  for builder in list(train_hook_builders or []) + (list(chief_train_hook_builders or []) if is_chief else []):
    if not hasattr(builder, 'create_hooks'):
      raise NotImplementedError('train hooks must come from a tensor2robot_b200.hooks.HookBuilder; TF SessionRunHooks '
                                'have no analogue here')
    hooks.extend(builder.create_hooks(t2r_model, model_dir))
  if t2r_model is None:
    raise ValueError('t2r_model is required')
  device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
  print_specification(t2r_model)
  result = {'global_step': t2r_model.global_step, 'loss': None, 'eval': None}
  run_config = t2r_model.get_run_config()
  if input_generator_train is not None:
    provide_input_generator_with_model_information(input_generator_train, t2r_model, ModeKeys.TRAIN)
    resumed = False
    started = False
    last_loss = None
    for features, labels in _batches(input_generator_train, t2r_model, ModeKeys.TRAIN, device):
      if not resumed:
        t2r_model.build(features, labels)
        # the chief decides what to resume from and every rank follows (a rank globbing model_dir on its own could
        # pick a checkpoint the chief is about to rotate away, or disagree about the starting step)
        ckpt = [latest_checkpoint(model_dir) if is_chief else None]
        if distributed:
          torch.distributed.broadcast_object_list(ckpt, src=0)
        if ckpt[0]:
          t2r_model.load_state_dict(torch.load(ckpt[0], weights_only=False))
        else:
          save_checkpoint(t2r_model, model_dir, run_config.get('keep_checkpoint_max', 5))   # model.ckpt-0
        if distributed:
          torch.distributed.barrier()
        resumed = True
      if t2r_model.global_step >= max_train_steps:
        break
      if not started:
        for hook in hooks:
          hook.begin()
        started = True
      for hook in hooks:
        hook.before_step(t2r_model.global_step)
      last_loss = t2r_model.train_step(features, labels)
      step = t2r_model.global_step
      for hook in hooks:
        hook.after_step(step, last_loss)
      if step % log_every_n_steps == 0:
        logging.info('step %d loss %.5f', step, float(last_loss))
      if step % run_config.get('save_checkpoints_steps', 1000) == 0:
        save_checkpoint(t2r_model, model_dir, run_config.get('keep_checkpoint_max', 5))
    if resumed:
      save_checkpoint(t2r_model, model_dir, run_config.get('keep_checkpoint_max', 5))
    if started:
      for hook in hooks:
        hook.end()
    result['global_step'] = t2r_model.global_step
    result['loss'] = float(last_loss) if last_loss is not None else None
  if input_generator_eval is not None:
    provide_input_generator_with_model_information(input_generator_eval, t2r_model, ModeKeys.EVAL)
    from tensor2robot_b200.utils import metrics as metrics_lib
    losses, streaming = [], metrics_lib.Accumulator()
    for i, (features, labels) in enumerate(_batches(input_generator_eval, t2r_model, ModeKeys.EVAL, device)):
      if eval_steps is not None and i >= eval_steps:
        break
      if not t2r_model.variable_store.finalized:
        t2r_model.build(features, labels)
      with t2r_model.averaged_parameters():   # the Estimator evaluates checkpoints = the swapped-in averages
        out = t2r_model.model_fn(features, labels, ModeKeys.EVAL)
      losses.append(out.loss.detach())
      streaming.update(out.eval_metrics)            # model_eval_fn's streaming metrics: statistics summed over the batches
    if losses:
      result['eval'] = dict(streaming.results(), loss=float(torch.stack(losses).mean()), steps=len(losses))
  return result


def predict_from_model(t2r_model=None, input_generator_predict=None, model_dir=None, device=None):
  """Yields predictions for every batch of the generator (:390-421)."""
  device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
  provide_input_generator_with_model_information(input_generator_predict, t2r_model, ModeKeys.PREDICT)
  loaded = False
  for features, _ in _batches(input_generator_predict, t2r_model, ModeKeys.PREDICT, device):
    if not loaded:
      t2r_model.build(features, mode=ModeKeys.PREDICT)
      ckpt = latest_checkpoint(model_dir) if model_dir else None
      if ckpt:
        t2r_model.load_state_dict(torch.load(ckpt, weights_only=False), restore_training_state=False)
      loaded = True
    yield t2r_model.predict(features)
