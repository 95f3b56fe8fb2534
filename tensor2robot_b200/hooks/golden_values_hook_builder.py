"""Golden-value regression harness (hooks/golden_values_hook_builder.py:34-79): models register the tensors worth
tracking with `add_golden_tensor(tensor, name)` while they build their losses; the hook records them after every train
step and writes `model_dir/golden_values.npy` (a list of {name: value} dicts) at the end of training."""
import logging
import os

import numpy as np
import torch

from tensor2robot_b200.hooks import hook_builder

COLLECTION = 'golden'
PREFIX = 'golden_'
_COLLECTION = {}          # name -> tensor of the current step (the TF graph collection)


def add_golden_tensor(tensor, name):
  """Adds a tensor to be tracked."""
  _COLLECTION[name] = tensor


def get_collection():
  return dict(_COLLECTION)


def clear_collection():
  _COLLECTION.clear()


class GoldenValuesHook(hook_builder.TrainHook):
  """Saves the registered values of every step to a file."""

  def __init__(self, log_directory):
    self._log_directory = log_directory
    self._measurements = []

  def begin(self):
    self._measurements = []

  def before_step(self, step):
    del step
    clear_collection()

  def after_step(self, step, loss):
    del step, loss
    golden_values = {}
    for name, value in _COLLECTION.items():
      golden_values[name] = value.detach().float().cpu().numpy() if torch.is_tensor(value) else np.asarray(value)
    logging.info('Recorded golden values for %s', golden_values.keys())
    self._measurements.append(golden_values)

  def end(self):
    os.makedirs(self._log_directory, exist_ok=True)
    np.save(os.path.join(self._log_directory, 'golden_values.npy'), np.array(self._measurements, dtype=object),
            allow_pickle=True)


class GoldenValuesHookBuilder(hook_builder.HookBuilder):
  """Hook builder for generating golden values."""

  def create_hooks(self, t2r_model, model_dir):
    del t2r_model
    return [GoldenValuesHook(model_dir)]
