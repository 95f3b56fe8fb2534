"""Export listeners of the training loop (behaviour of the reference's hooks/checkpoint_hooks.py:28-201,
pinned by the scenarios of its checkpoint_hooks_test.py, which tests/test_hooks.py replays).

Two contracts:

* `CheckpointExportListener.after_save(session, global_step)` calls `export_fn(export_dir, global_step)`, which writes
  one version directory and returns its path; with `num_versions` only that many newest versions are kept.
* `LaggedCheckpointListener` additionally maintains `lagged_export_dir` so that its newest version is always the
  export that was current BEFORE the latest one (the very first export is mirrored immediately, so the directory is
  never empty): the lagged / target network that TD3 and QT-Opt actors and Bellman updaters poll.  The in-process,
  on-device equivalent is `engine.LaggedTarget`.

Both are written as operations on a `_VersionDir` (an ordered set of version directories with a retention limit);
the lagged listener is one idempotent rule - "make `name` the newest lagged version" - applied at construction (which
is what repairs or resumes from directories left by an earlier run) and after every export."""
import logging
import os
import re
import shutil


def _natural_key(name):
  return [int(tok) if tok.isdigit() else tok for tok in re.split(r'(\d+)', name)]


class _VersionDir(object):
  """Version directories under `root`, oldest first: what was on disk at construction (sorted), then every `add`."""

  def __init__(self, root, keep=None):
    self.root = str(root)
    self.keep = keep
    os.makedirs(self.root, exist_ok=True)
    self.versions = sorted(os.listdir(self.root), key=_natural_key)
    self._prune()

  def path(self, name):
    return os.path.join(self.root, name)

  @property
  def newest(self):
    return self.versions[-1] if self.versions else None

  def add(self, name):
    if name in self.versions:
      self.versions.remove(name)
    self.versions.append(name)
    self._prune()

  def _prune(self):
    while self.keep and len(self.versions) > self.keep:
      shutil.rmtree(self.path(self.versions.pop(0)), ignore_errors=True)


class CheckpointExportListener(object):
  """Exports the model after a checkpoint was written; keeps the `num_versions` newest exports (None: all)."""

  def __init__(self, export_fn, export_dir, num_versions=None):
    self._export_fn = export_fn
    self._exports = _VersionDir(export_dir, num_versions)

  def _export(self, global_step):
    path = str(self._export_fn(self._exports.root, global_step))
    logging.info('exported the model of global step %d to %s', global_step, path)
    return path, os.path.basename(os.path.normpath(path))

  def after_save(self, session, global_step):
    del session
    path, name = self._export(global_step)
    self._exports.add(name)
    return path


class LaggedCheckpointListener(CheckpointExportListener):
  """Keeps `lagged_export_dir` one export behind `export_dir`."""

  def __init__(self, export_fn, export_dir, lagged_export_dir, num_versions):
    super(LaggedCheckpointListener, self).__init__(export_fn, export_dir, num_versions)
    self._lagged = _VersionDir(lagged_export_dir, num_versions)
    on_disk = self._exports.versions
    if on_disk:   # the version before the newest, or the only one
      self._make_newest_lagged(on_disk[-2] if len(on_disk) > 1 else on_disk[-1])

  def _make_newest_lagged(self, name):
    if self._lagged.newest == name:
      return
    shutil.copytree(self._exports.path(name), self._lagged.path(name), dirs_exist_ok=True)
    self._lagged.add(name)

  def after_save(self, session, global_step):
    del session
    previous = self._exports.newest
    path, name = self._export(global_step)
    self._make_newest_lagged(previous if previous is not None else name)   # before retention can retire `previous`
    self._exports.add(name)
    return path
