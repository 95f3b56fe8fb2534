"""AbstractT2RModel on the B200 engine.

Keeps the reference's model decomposition (models/abstract_model.py:162-981):
  get_feature_specification / get_label_specification (no batch dim)  -> preprocessor
  inference_network_fn(features, labels, mode, config, params) -> dict | (dict, update_ops)   :405-451
  model_train_fn(features, labels, inference_outputs, mode, ...) -> loss | (loss, dict)       :454-504
  model_eval_fn(...) -> metrics dict                                                           :506-565
  create_optimizer(params) / create_train_op(...)                                              :335-381, :836-871
  model_fn(features, labels, mode, config, params)                                             :662-834
and replaces what TensorFlow did underneath: there is no graph / EstimatorSpec.  `model_fn` executes
eagerly on the current CUDA stream inside the model's VariableStore and returns a `ModelOutput`;
`train_step` = model_fn(TRAIN) + backward + (all-reduce) + fused optimizer kernel, i.e. one
Session.run of the reference's train_op.
"""
import abc
import collections
import contextlib
import logging

import torch
import torch.distributed as dist

from tensor2robot_b200 import nn
from tensor2robot_b200.models import model_interface
from tensor2robot_b200.models import optimizers
from tensor2robot_b200.utils import tensorspec_utils

TRAIN, EVAL, PREDICT = model_interface.TRAIN, model_interface.EVAL, model_interface.PREDICT
PCGRAD_LOSSES_COLLECTION = 'pcgrad_losses'   # key of the task-loss list in model_train_fn's train_outputs
DEVICE_TYPE_CPU, DEVICE_TYPE_GPU, DEVICE_TYPE_TPU = 'cpu', 'gpu', 'tpu'

ModelOutput = collections.namedtuple('ModelOutput', ['mode', 'loss', 'train_outputs', 'predictions', 'eval_metrics'])


class _RestorableVariable(object):
  """What filter_restorables_fn sees: `.name` / `.op.name` = the reference variable name, `.shape`."""

  def __init__(self, name, shape):
    self.name = name
    self.shape = tuple(shape)
    self.op = self


def default_init_from_checkpoint_fn(checkpoint, allow_partial_restore=False, filter_restorables_fn=None):
  """init_from_checkpoint_fn that initialises a model from a TensorFlow (tensor-bundle) checkpoint by variable name
  (abstract_model.py:87-126).  Returns the callable the model invokes once its variables exist.

  allow_partial_restore: tolerate model variables that are missing in the checkpoint (otherwise ValueError).
  filter_restorables_fn: optional predicate on a variable (`.name`, `.op.name`, `.shape`) choosing what to restore."""
  from tensor2robot_b200.utils import tf_checkpoint

  def init_fn(t2r_model):
    logging.info('Initializing model weights from %s', checkpoint)
    reader = tf_checkpoint.load_checkpoint(checkpoint)
    vs = t2r_model.variable_store
    arrays = {}
    for name, value in vs.export_tf().items():
      if filter_restorables_fn is not None and not filter_restorables_fn(_RestorableVariable(name, value.shape)):
        continue
      if reader.has_tensor(name):
        logging.info('Loading variable %s from checkpoint', name)
        arrays[name] = reader.get_tensor(name)
      elif allow_partial_restore:
        logging.warning('Variable %s is not in the checkpoint, skipping.', name)
      else:
        raise ValueError('Attempting to restore variable {} which is not in the checkpoint.'.format(name))
    vs.import_tf(arrays, strict=False)

  return init_fn


class AbstractT2RModel(model_interface.ModelInterface):
  """Base class encapsulating a model_fn and metadata about input/output sizes."""

  def __init__(self, preprocessor_cls=None, create_optimizer_fn=optimizers.default_create_optimizer_fn,
               device_type=DEVICE_TYPE_GPU, summarize_gradients=True, use_sync_replicas_optimizer=False,
               use_avg_model_params=False, init_from_checkpoint_fn=None, device=None, seed=0):
    self._preprocessor_cls = preprocessor_cls
    self._create_optimizer_fn = create_optimizer_fn
    self._device_type = device_type
    self._summarize_gradients = summarize_gradients
    self._use_sync_replicas_optimizer = use_sync_replicas_optimizer
    self._use_avg_model_params = use_avg_model_params
    self._init_from_checkpoint_fn = init_from_checkpoint_fn
    self._optimizer = None
    self._reducer = None
    self._device = device
    self._seed = seed
    self._vs = None
    self.global_step = 0

  # -- engine state ---------------------------------------------------------------------------
  @property
  def variable_store(self):
    """The model's variables (created lazily on the model's CUDA device)."""
    if self._vs is None:
      device = self._device or torch.device('cuda', torch.cuda.current_device())
      self._vs = nn.VariableStore(device, seed=self._seed)
    return self._vs

  def l2_regularization(self):
    """slim l2_regularizer scale applied to the regularised weights (0 = none)."""
    return 0.0

  # -- specs / preprocessor (abstract_model.py:220-276) -----------------------------------------
  def create_pack_features(self, feature_spec, label_spec):
    raise NotImplementedError()

  def get_feature_specification_for_packing(self, mode):
    return self.preprocessor.get_in_feature_specification(mode)

  def get_label_specification_for_packing(self, mode):
    return self.preprocessor.get_in_label_specification(mode)

  @property
  def default_preprocessor_cls(self):
    from tensor2robot_b200.preprocessors import noop_preprocessor
    return noop_preprocessor.NoOpPreprocessor

  @property
  def preprocessor(self):
    preprocessor_cls = self._preprocessor_cls
    if preprocessor_cls is None:
      preprocessor_cls = self.default_preprocessor_cls
    return preprocessor_cls(model_feature_specification_fn=self.get_feature_specification,
                            model_label_specification_fn=self.get_label_specification,
                            is_model_device_tpu=self.is_device_tpu)

  @abc.abstractmethod
  def get_feature_specification(self, mode):
    """Required features for the model_fn / inference_network_fn (no batch dimension)."""

  @abc.abstractmethod
  def get_label_specification(self, mode):
    """Required labels for the model_fn / model_train_fn / model_eval_fn."""

  # -- the model decomposition ------------------------------------------------------------------
  @abc.abstractmethod
  def inference_network_fn(self, features, labels, mode, config=None, params=None):
    """The inference network; returns a dict of outputs or (dict, update_ops)."""

  @abc.abstractmethod
  def model_train_fn(self, features, labels, inference_outputs, mode, config=None, params=None):
    """The training loss; returns loss or (loss, train_outputs dict)."""

  def model_eval_fn(self, features, labels, inference_outputs, train_loss, train_outputs, mode, config=None,
                    params=None):
    """Evaluation metrics (abstract_model.py:506-565): by default the loss."""
    del features, labels, inference_outputs, train_outputs, mode, config, params
    return {'loss': train_loss}

  def add_summaries(self, features, labels, inference_outputs, train_loss, train_outputs, mode, config=None,
                    params=None):
    del features, labels, inference_outputs, train_loss, train_outputs, mode, config, params

  def create_export_outputs_fn(self, features, inference_outputs, mode, config=None, params=None):
    """Predictions exposed by predictors (abstract_model.py:610-660): the inference outputs."""
    del features, mode, config, params
    return inference_outputs

  def create_optimizer(self, params=None):
    """abstract_model.py:836-871; sync-replica / tower wrappers are replaced by the single NCCL
    all-reduce in train_step."""
    optimizer = self._create_optimizer_fn(self.use_summaries(params))
    if self._use_avg_model_params and not isinstance(optimizer, optimizers.MovingAverageOptimizer):
      optimizer = optimizers.create_moving_average_optimizer(optimizer)
    optimizer.l2_regularization = self.l2_regularization()
    return optimizer

  @property
  def optimizer(self):
    if self._optimizer is None:
      self._optimizer = self.create_optimizer()
    return self._optimizer

  def use_summaries(self, params=None):
    if params is None:
      return True
    return params.get('use_summaries', True) if isinstance(params, dict) else True

  def maybe_init_from_checkpoint(self):
    if self._init_from_checkpoint_fn is not None:
      self._init_from_checkpoint_fn(self)

  # -- model_fn (abstract_model.py:662-834) -----------------------------------------------------
  def model_fn(self, features, labels, mode, config=None, params=None):
    feature_spec = self.get_feature_specification(mode)
    with tensorspec_utils.device_bf16_as_float32():
      features = tensorspec_utils.validate_and_pack(feature_spec, features, ignore_batch=True)
      if labels is not None:
        labels = tensorspec_utils.validate_and_pack(self.get_label_specification(mode), labels, ignore_batch=True)
    vs = self.variable_store
    grad = torch.enable_grad() if mode == TRAIN else torch.no_grad()
    with grad, nn.variable_store(vs):
      inference_outputs = self.inference_network_fn(features, labels, mode, config, params)
      if isinstance(inference_outputs, tuple):
        inference_outputs = inference_outputs[0]   # update ops run inside the fused BN kernels
      if mode == PREDICT:
        predictions = self.create_export_outputs_fn(features, inference_outputs, mode, config, params)
        return ModelOutput(mode, None, None, predictions, None)
      train_fn_result = self.model_train_fn(features, labels, inference_outputs, mode, config, params)
      if isinstance(train_fn_result, tuple):
        train_loss, train_outputs = train_fn_result
      else:
        train_loss, train_outputs = train_fn_result, {}
      if mode == TRAIN:
        return ModelOutput(mode, train_loss, train_outputs, inference_outputs, None)
      metrics = self.model_eval_fn(features, labels, inference_outputs, train_loss, train_outputs, mode, config,
                                   params)
      return ModelOutput(mode, train_loss, train_outputs, inference_outputs, metrics)

  def build(self, features, labels=None, mode=EVAL):
    """Creates the variables with an inference pass over the first two examples and consolidates
    them into the flat parameter / gradient buffers (nn.VariableStore.finalize)."""
    vs = self.variable_store
    if vs.finalized:
      return
    def first_two(struct):
      return tensorspec_utils.TensorSpecStruct(
          [(k, v[:2]) for k, v in tensorspec_utils.flatten_spec_structure(struct).items()])

    # An inference-mode pass creates every variable (graph structure does not depend on the mode;
    # batch norm merely reads its moving statistics instead of updating them).
    with tensorspec_utils.device_bf16_as_float32():
      head = tensorspec_utils.validate_and_pack(self.get_feature_specification(mode), first_two(features),
                                                ignore_batch=True)
    with torch.no_grad(), nn.variable_store(vs):
      self.inference_network_fn(head, first_two(labels) if labels is not None else None, mode, None, None)
    vs.finalize()
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if world > 1:
      dist.broadcast(vs.flat, src=0)
      dist.broadcast(vs.state_flat, src=0)
      vs.sync_compute_copies()
    self.maybe_init_from_checkpoint()

  def _build_mode_is_predict(self):
    """PREDICT feature specs may differ from TRAIN ones (tiled actions): build with EVAL-shaped
    features unless a subclass says the PREDICT graph has the same variables."""
    return False

  # -- one optimisation step (the reference's train_op: create_train_op :335-381) ---------------
  def train_step(self, features, labels, config=None, params=None):
    """features/labels: (flat or packed) structures of batched CUDA tensors as produced by the
    preprocessor.  Returns the device scalar loss; never synchronises."""
    vs = self.variable_store
    if not vs.finalized:
      self.build(features, labels)
    out = self.model_fn(features, labels, TRAIN, config, params)
    task_losses = (out.train_outputs or {}).get(PCGRAD_LOSSES_COLLECTION) if isinstance(out.train_outputs, dict) else None
    from tensor2robot_b200 import engine   # the shared step tail (backward, bucketed all-reduce, fused optimizer)
    if self._reducer is None or self._reducer.vs is not vs:
      self._reducer = engine.GradientReducer(vs)
    engine.optimization_step(vs, self.optimizer, self.global_step, out.loss, task_losses, self._reducer)
    self.global_step += 1
    total = out.loss.detach()
    l2 = self.l2_regularization()
    if l2:
      total = total + nn.l2_regularization_loss(l2, vs)
    return total

  def predict(self, features, config=None, params=None, high_precision=False):
    """PREDICT outputs.  high_precision=True evaluates the graph in nn.high_precision() (fp32 activations, bf16x3
    convolutions): what serving / CEM callers that need Q values within 1e-3 relative of an fp32 evaluation ask for."""
    if not high_precision:
      return self.model_fn(features, None, PREDICT, config, params).predictions
    with torch.no_grad(), nn.high_precision():
      return self.model_fn(features, None, PREDICT, config, params).predictions

  # -- checkpoints (TF variable names / layouts) --------------------------------------
  # With a MovingAverageOptimizer (use_avg_model_params) the reference installs optimizer.swapping_saver
  # (models/abstract_model.py:855-863): checkpoints and exports hold the AVERAGED parameters under the variable
  # names, which is what predictors, the lagged target export and evaluation load; the raw parameters travel beside
  # them so that training resumes exactly.
  def _moving_average_shadow(self):
    return optimizers.moving_average_shadow(self._optimizer)

  @contextlib.contextmanager
  def averaged_parameters(self):
    """Runs the body with the averaged parameters swapped in (no-op without a moving average)."""
    vs, ema = self.variable_store, self._moving_average_shadow()
    if ema is None or not vs.finalized:
      yield
      return
    raw = vs.flat.clone()
    vs.flat.copy_(ema)
    vs.sync_compute_copies()
    try:
      yield
    finally:
      vs.flat.copy_(raw)
      vs.sync_compute_copies()

  def export_variables(self):
    """{reference variable name: array in the TF layout}, averaged parameters when a moving average exists."""
    vs, ema = self.variable_store, self._moving_average_shadow()
    if ema is None or not vs.finalized:
      return vs.export_tf()
    raw = vs.flat.clone()
    vs.flat.copy_(ema)
    try:
      return vs.export_tf()
    finally:
      vs.flat.copy_(raw)

  def state_dict(self):
    vs = self.variable_store
    state = {'variables': self.export_variables(), 'optimizer': self.optimizer.state_dict() if self._optimizer else {},
             'global_step': self.global_step}
    if self._moving_average_shadow() is not None and vs.finalized:
      state['raw_parameters'] = vs.flat.detach().cpu()    # 'variables' are the averages
    return state

  def load_state_dict(self, state, restore_training_state=True):
    """restore_training_state=False (serving): keep whatever 'variables' holds - the averaged parameters when the
    checkpoint was written with a moving average - and ignore the optimizer slots."""
    vs = self.variable_store
    vs.import_tf(state['variables'])
    if restore_training_state:
      if state.get('optimizer'):
        self.optimizer.load_state_dict(state['optimizer'], vs)
      raw = state.get('raw_parameters')
      if raw is not None and raw.numel() == vs.flat.numel():
        vs.flat.copy_(raw.to(vs.flat.device))
        vs.sync_compute_copies()
    self.global_step = int(state.get('global_step', 0))

  # -- run config / device ----------------------------------------------------------------------
  def get_run_config(self):
    return {'save_checkpoints_steps': 1000, 'keep_checkpoint_max': 5}

  @property
  def is_device_tpu(self):
    return self._device_type == DEVICE_TYPE_TPU

  @property
  def is_device_gpu(self):
    return self._device_type == DEVICE_TYPE_GPU

  @property
  def device_type(self):
    return self._device_type

  @device_type.setter
  def device_type(self, device_type):
    self._device_type = device_type
