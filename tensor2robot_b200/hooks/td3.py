"""TD3Hooks (hooks/td3.py:37-132): while training, export the latest model into `export_dir` and a version lagged by one
export interval into `lagged_export_dir` (Fujimoto et al., "Addressing Function Approximation Error in Actor-Critic
Methods"): the target network of distributed QT-Opt / TD3 that actors and Bellman updaters poll.

The reference exports SavedModels every `save_secs`; here an export is `<dir>/<global_step>/` with the model variables as
a TensorFlow-bundle checkpoint (reference variable names and layouts, `utils/tf_checkpoint.py`) plus
`assets.extra/t2r_assets.pbtxt` (the feature / label specs and the global step), written every `save_steps` optimizer
steps (wall-clock triggers would make runs irreproducible)."""
import os

from tensor2robot_b200.hooks import checkpoint_hooks
from tensor2robot_b200.hooks import hook_builder
from tensor2robot_b200.proto import t2r_pb2
from tensor2robot_b200.utils import tensorspec_utils

PREDICT = 'infer'


def export_model(t2r_model, export_dir, global_step):
  """Writes `<export_dir>/<global_step>/{model.ckpt-*, checkpoint, assets.extra/t2r_assets.pbtxt}`; returns the path."""
  from tensor2robot_b200.utils import train_eval
  path = os.path.join(export_dir, '%010d' % int(global_step))
  os.makedirs(os.path.join(path, 'assets.extra'), exist_ok=True)
  train_eval.save_tf_checkpoint(t2r_model, path)
  t2r_assets = t2r_pb2.T2RAssets()
  t2r_assets.feature_spec.CopyFrom(t2r_model.get_feature_specification_for_packing(PREDICT).to_proto())
  t2r_assets.label_spec.CopyFrom(t2r_model.get_label_specification_for_packing(PREDICT).to_proto())
  t2r_assets.global_step = int(global_step)
  tensorspec_utils.write_t2r_assets_to_file(t2r_assets, os.path.join(path, 'assets.extra',
                                                                     tensorspec_utils.T2R_ASSETS_FILENAME))
  return path


class _ExportHook(hook_builder.TrainHook):

  def __init__(self, listener, save_steps):
    self._listener = listener
    self._save_steps = save_steps
    self._last = None

  def after_step(self, step, loss):
    del loss
    if step % self._save_steps == 0 and step != self._last:
      self._listener.after_save(None, step)
      self._last = step


class TD3Hooks(hook_builder.HookBuilder):
  """export_dir: latest models; lagged_export_dir: models lagged by one export; num_versions kept in each."""

  def __init__(self, export_dir, lagged_export_dir, batch_sizes_for_export=None, save_steps=1000, num_versions=3,
               export_fn=None):
    del batch_sizes_for_export          # TF-Serving warm-up requests have no analogue
    self._export_dir = export_dir
    self._lagged_export_dir = lagged_export_dir
    self._save_steps = save_steps
    self._num_versions = num_versions
    self._export_fn = export_fn

  def create_hooks(self, t2r_model, model_dir):
    del model_dir
    if not self._export_dir and not self._lagged_export_dir:
      return []
    export_fn = self._export_fn or (lambda export_dir, global_step: export_model(t2r_model, export_dir, global_step))
    listener = checkpoint_hooks.LaggedCheckpointListener(export_fn=export_fn, num_versions=self._num_versions,
                                                         export_dir=self._export_dir,
                                                         lagged_export_dir=self._lagged_export_dir)
    return [_ExportHook(listener, self._save_steps)]
