"""Export listeners (hooks/checkpoint_hooks.py:28-201): after a checkpoint, export the model into `export_dir`; the
lagged listener additionally keeps `lagged_export_dir` exactly one export behind - the TD3 / QT-Opt target network that
remote actors and Bellman updaters load.  An export here is a directory
`<export_dir>/<global_step>/` holding a TensorFlow-bundle checkpoint + `assets.extra/t2r_assets.pbtxt`
(`export_fn` decides); the in-process, on-device equivalent is `engine.LaggedTarget`."""
import collections
import logging
import os
import shutil


class _DirectoryVersionGC(object):
  """Observes a stream of incoming directories and removes the oldest ones."""

  def __init__(self, num_versions):
    self._queue = collections.deque()
    self._num_versions = num_versions

  def observe(self, directory):
    self._queue.append(directory)
    self._remove_if_necessary()

  def observe_multiple(self, directory_list):
    self._queue.extend(directory_list)
    self._remove_if_necessary()

  def _remove_if_necessary(self):
    while len(self._queue) > self._num_versions:
      shutil.rmtree(self._queue.popleft(), ignore_errors=True)


class CheckpointExportListener(object):
  """Exports the model after a checkpoint was created."""

  def __init__(self, export_fn, export_dir, num_versions=None):
    """export_fn(export_dir, global_step) -> exported path; num_versions: exports to keep (None: all)."""
    self._export_fn = export_fn
    self._export_dir = str(export_dir)
    os.makedirs(self._export_dir, exist_ok=True)
    self._gc = None
    if num_versions:
      self._gc = _DirectoryVersionGC(num_versions)
      self._gc.observe_multiple([os.path.join(self._export_dir, f) for f in sorted(os.listdir(self._export_dir))])

  def after_save(self, session, global_step):
    del session
    logging.info('Exporting model at global_step %d', global_step)
    exported_path = str(self._export_fn(self._export_dir, global_step))
    logging.info('Saved model to %s', exported_path)
    if self._gc:
      self._gc.observe(exported_path)
    return exported_path


class LaggedCheckpointListener(CheckpointExportListener):
  """Also exports the **second newest** model to a separate directory (the lagged / target network)."""

  def __init__(self, export_fn, export_dir, lagged_export_dir, num_versions):
    CheckpointExportListener.__init__(self, export_fn, export_dir, num_versions)
    self._lagged_export_dir = str(lagged_export_dir)
    self._current_model_dir = None
    self._lagged_model_dir = None
    self._lagged_gc = _DirectoryVersionGC(num_versions) if self._gc else None
    os.makedirs(self._lagged_export_dir, exist_ok=True)
    export_dir_contents = sorted(os.listdir(self._export_dir))
    lagged_export_dir_contents = sorted(os.listdir(self._lagged_export_dir))
    if self._lagged_gc:
      self._lagged_gc.observe_multiple([os.path.join(self._lagged_export_dir, f) for f in lagged_export_dir_contents])
    # resume: re-establish "lagged is one export behind current" from what is on disk (:137-160)
    if len(export_dir_contents) == 1:
      self._current_model_dir = os.path.join(self._export_dir, export_dir_contents[0])
      if export_dir_contents == lagged_export_dir_contents:
        self._lagged_model_dir = os.path.join(self._lagged_export_dir, lagged_export_dir_contents[0])
      else:
        self._lagged_model_dir = self._copy_savedmodel(self._current_model_dir, self._lagged_export_dir)
    elif len(export_dir_contents) > 1:
      second_last_exported_model = export_dir_contents[-2]
      self._current_model_dir = os.path.join(self._export_dir, export_dir_contents[-1])
      if not lagged_export_dir_contents or second_last_exported_model != lagged_export_dir_contents[-1]:
        self._lagged_model_dir = self._copy_savedmodel(os.path.join(self._export_dir, second_last_exported_model),
                                                       self._lagged_export_dir)
      else:
        self._lagged_model_dir = os.path.join(self._lagged_export_dir, lagged_export_dir_contents[-1])

  def _copy_savedmodel(self, source_dir, destination):
    dest_base_dir = os.path.join(str(destination), os.path.basename(str(source_dir)))
    shutil.copytree(str(source_dir), dest_base_dir, dirs_exist_ok=True)
    return dest_base_dir

  def _copy_lagged_model(self, source_dir, destination):
    destination_path = self._copy_savedmodel(source_dir, destination)
    if self._lagged_gc:
      self._lagged_gc.observe(destination_path)
    return destination_path

  def after_save(self, session, global_step):
    """Exports to the current directory; the lagged directory receives the export that was current until now."""
    export_dir = CheckpointExportListener.after_save(self, session, global_step)
    if not self._current_model_dir:
      self._lagged_model_dir = self._copy_lagged_model(export_dir, self._lagged_export_dir)
    elif os.path.basename(self._current_model_dir) == os.path.basename(self._lagged_model_dir):
      pass
    else:
      self._lagged_model_dir = self._copy_lagged_model(self._current_model_dir, self._lagged_export_dir)
    self._current_model_dir = export_dir
    return export_dir
