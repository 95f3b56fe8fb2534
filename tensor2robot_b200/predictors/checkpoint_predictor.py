"""CheckpointPredictor: a predictor over the engine's checkpoints (predictors/checkpoint_predictor.py:37-215).

The reference builds a PREDICT graph with placeholders and a session; here the model's variables live in its
VariableStore on the GPU and predict() is: numpy features -> pinned staging + H2D -> the model's preprocessor ->
model.predict (the CUDA kernels) -> numpy predictions.  There is no CPU path (`use_gpu=False` is rejected)."""
import logging
import os
import time

import numpy as np
import torch

from tensor2robot_b200 import nn
from tensor2robot_b200.predictors import abstract_predictor
from tensor2robot_b200.utils import tensorspec_utils
from tensor2robot_b200.utils import train_eval

PREDICT = 'infer'
_BUSY_WAITING_SLEEP_TIME_IN_SECS = 1


class CheckpointPredictor(abstract_predictor.AbstractPredictor):

  def __init__(self, t2r_model, checkpoint_dir=None, use_gpu=True, timeout=600, device=None, high_precision=False):
    if not use_gpu:
      raise ValueError('the B200 engine has no CPU inference path (use_gpu=False)')
    self._checkpoint_dir = checkpoint_dir
    self._timeout = timeout
    self._t2r_model = t2r_model
    self._high_precision = high_precision   # nn.high_precision(): fp32 activations, bf16x3 convolutions
    self._preprocessor = t2r_model.preprocessor
    feature_tspec = self._preprocessor.get_in_feature_specification(PREDICT)
    # inference: only the required tensors
    self._feature_tspec = tensorspec_utils.filter_required_flat_tensor_spec(feature_tspec)
    self._label_tspec = self._feature_tspec     # the reference exposes the in-feature spec here as well (:82-84)
    self._device = device
    self._stager = None
    self._current_checkpoint_path = None
    self._model_was_restored = False

  # -- device plumbing ----------------------------------------------------------------------------
  def _dev(self):
    if self._device is None:
      self._device = torch.device('cuda', torch.cuda.current_device())
    return torch.device(self._device)

  def _preprocess(self, features):
    device = self._dev()
    if self._stager is None:
      self._stager = train_eval.DeviceStager(device)
    flat = tensorspec_utils.flatten_spec_structure(features)
    staged, ready = self._stager.stage(tensorspec_utils.TensorSpecStruct(
        [(k, np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in flat.items()]))
    torch.cuda.current_stream(device).wait_event(ready)
    processed, _ = self._preprocessor.preprocess(staged, None, PREDICT)
    return processed

  def _ensure_built(self):
    if not self._t2r_model.variable_store.finalized:
      random_features = tensorspec_utils.make_random_numpy(self._feature_tspec, batch_size=2)
      self._t2r_model.build(self._preprocess(random_features), mode=PREDICT)

  # -- AbstractPredictor --------------------------------------------------------------------------
  def predict(self, features):
    self.assert_is_loaded()
    with torch.no_grad():
      if self._high_precision:
        with nn.high_precision():          # the preprocessor sees the mode too and keeps images in fp32
          predictions = self._t2r_model.predict(self._preprocess(features), high_precision=True)
      else:
        predictions = self._t2r_model.predict(self._preprocess(features))
    return {k: v.detach().float().cpu().numpy() if torch.is_tensor(v) else v for k, v in predictions.items()}

  def get_feature_specification(self):
    return self._feature_tspec

  def get_label_specification(self):
    return self._label_tspec

  def init_randomly(self):
    logging.info('Initializing model with random weights')
    self._ensure_built()
    self._model_was_restored = True

  def restore(self):
    """True if a (new or unchanged) checkpoint is loaded, False if none appeared within `timeout` seconds."""
    if self._checkpoint_dir is None:
      raise ValueError('The predictor cannot be restored since no checkpoint_dir has been passed.')
    if '.ckpt-' in os.path.basename(self._checkpoint_dir):
      latest = self._checkpoint_dir
    else:
      start_time = time.time()
      latest = None
      while time.time() - start_time < self._timeout and latest is None:
        latest = train_eval.latest_checkpoint(self._checkpoint_dir)
        if latest is None:
          logging.warning('No checkpoint found at %s:\nThe next attempt to check for latest model will be in %d seconds',
                          self._checkpoint_dir, _BUSY_WAITING_SLEEP_TIME_IN_SECS)
          time.sleep(_BUSY_WAITING_SLEEP_TIME_IN_SECS)
      if latest is None:
        return False
    if latest == self._current_checkpoint_path:
      logging.info("Checkpoint '%s' wasn't updated.", latest)
      return True
    self._ensure_built()
    state = torch.load(latest, weights_only=False)
    # the PREDICT side holds only the model variables, not the optimizer slots (:100-104)
    self._t2r_model.load_state_dict({'variables': state['variables'], 'global_step': state.get('global_step', 0)})
    self._current_checkpoint_path = latest
    self._model_was_restored = True
    return True

  def close(self):
    self._stager = None
    self._model_was_restored = False

  def assert_is_loaded(self):
    if not self._model_was_restored:
      raise ValueError('The predictor has not yet been successfully restored.')

  @property
  def model_version(self):
    return self.global_step

  @property
  def global_step(self):
    try:
      self.assert_is_loaded()
    except ValueError:
      return -1
    return int(self._t2r_model.global_step)

  @property
  def model_path(self):
    self.assert_is_loaded()
    return self._current_checkpoint_path
