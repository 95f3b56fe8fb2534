"""ExportedModelPredictor: the counterpart of ExportedSavedModelPredictor (predictors/exported_savedmodel_predictor.py:
52-260) for the exports `hooks/td3.py` writes: `<export_dir>/<numbered version>/` holding a TensorFlow-bundle checkpoint
and `assets.extra/t2r_assets.pbtxt`.  `restore()` waits for / picks the newest version (lexicographic order), reads the
feature / label specs and the global step from the assets file and loads the variables into the model by name.  A
SavedModel carries its graph; these exports carry weights only, so the predictor is given the T2R model to run."""
import logging
import os
import time

import numpy as np

from tensor2robot_b200.predictors import checkpoint_predictor
from tensor2robot_b200.utils import tensorspec_utils
from tensor2robot_b200.utils import tf_checkpoint

_BUSY_WAITING_SLEEP_TIME_IN_SECS = 1


def valid_export_versions(export_dir):
  """Sorted version directories of export_dir that are complete (assets + a checkpoint index)."""
  if not os.path.isdir(export_dir):
    return []
  out = []
  for name in sorted(os.listdir(export_dir)):
    path = os.path.join(export_dir, name)
    if not (name.isdigit() and os.path.isdir(path)):
      continue           # temp-* directories of exports in flight, stray files
    assets = os.path.join(path, 'assets.extra', tensorspec_utils.T2R_ASSETS_FILENAME)
    if os.path.exists(assets) and any(f.endswith('.index') for f in os.listdir(path)):
      out.append(path)
  return out


class ExportedModelPredictor(checkpoint_predictor.CheckpointPredictor):

  def __init__(self, export_dir, t2r_model, timeout=600, device=None, high_precision=False):
    super(ExportedModelPredictor, self).__init__(t2r_model=t2r_model, checkpoint_dir=None, timeout=timeout, device=device,
                                                 high_precision=high_precision)
    self._export_dir = export_dir
    self._latest_export_dir = None
    self._exported_feature_spec = None
    self._exported_label_spec = None
    self._global_step = -1

  def predict(self, features):
    """Features whose shape equals the un-batched spec shape get a batch dimension (:103-114)."""
    self.assert_is_loaded()
    flat_spec = tensorspec_utils.flatten_spec_structure(self.get_feature_specification())
    expanded = {}
    for key, value in tensorspec_utils.flatten_spec_structure(features).items():
      spec = flat_spec[key] if key in flat_spec.keys() else None
      value = np.asarray(value)
      if spec is not None and tuple(spec.shape) == tuple(value.shape):
        value = np.expand_dims(value, 0)
      expanded[key] = value
    return super(ExportedModelPredictor, self).predict(expanded)

  def get_feature_specification(self):
    self.assert_is_loaded()
    return self._exported_feature_spec

  def get_label_specification(self):
    self.assert_is_loaded()
    return self._exported_label_spec

  def _newest_version(self):
    if os.path.isdir(self._export_dir) and os.path.basename(os.path.normpath(self._export_dir)).isdigit() and \
        os.path.isdir(os.path.join(self._export_dir, 'assets.extra')):
      return self._export_dir                    # export_dir points at one specific version
    versions = valid_export_versions(self._export_dir)
    return versions[-1] if versions else None

  def _load_version(self, path):
    """Reads assets + weights of one export into the model."""
    assets = tensorspec_utils.load_t2r_assets_to_file(
        os.path.join(path, 'assets.extra', tensorspec_utils.T2R_ASSETS_FILENAME))
    self._exported_feature_spec = tensorspec_utils.TensorSpecStruct.from_proto(assets.feature_spec)
    self._exported_label_spec = tensorspec_utils.TensorSpecStruct.from_proto(assets.label_spec)
    self._global_step = int(assets.global_step)
    reader = tf_checkpoint.load_checkpoint(path)
    self._ensure_built()
    vs = self._t2r_model.variable_store
    arrays = {}
    for name in vs.export_tf():
      if not reader.has_tensor(name):
        raise ValueError('export %s lacks variable %s' % (path, name))
      arrays[name] = reader.get_tensor(name)
    vs.import_tf(arrays)
    self._t2r_model.global_step = self._global_step

  def restore(self, is_async=False):
    """True once the newest export is loaded (or already was); False if none appeared within `timeout` seconds."""
    del is_async            # loading is a few host-to-device copies; no background thread
    start_time = time.time()
    newest = self._newest_version()
    while newest is None and time.time() - start_time < self._timeout:
      logging.warning('No export found at %s; next attempt in %d seconds', self._export_dir,
                      _BUSY_WAITING_SLEEP_TIME_IN_SECS)
      time.sleep(_BUSY_WAITING_SLEEP_TIME_IN_SECS)
      newest = self._newest_version()
    if newest is None:
      return False
    if newest == self._latest_export_dir:
      return True
    self._load_version(newest)
    self._latest_export_dir = newest
    self._current_checkpoint_path = newest
    self._model_was_restored = True
    return True

  @property
  def global_step(self):
    return self._global_step if self._model_was_restored else -1

  @property
  def model_version(self):
    if not self._model_was_restored:
      return -1
    return int(os.path.basename(os.path.normpath(self._latest_export_dir)))

  @property
  def model_path(self):
    self.assert_is_loaded()
    return self._latest_export_dir
