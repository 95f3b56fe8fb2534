"""HookBuilder: creates training hooks for a model (hooks/hook_builder.py:27-48).  The Estimator's SessionRunHook
protocol has no analogue in the stream-driven trainer; a hook here is an object with the optional methods
`begin()`, `before_step(step)`, `after_step(step, loss)` and `end()`, driven by train_eval.train_eval_model."""
import abc


class TrainHook(object):
  """No-op base class of the hooks train_eval_model drives."""

  def begin(self):
    pass

  def before_step(self, step):
    del step

  def after_step(self, step, loss):
    del step, loss

  def end(self):
    pass


class HookBuilder(abc.ABC):

  @abc.abstractmethod
  def create_hooks(self, t2r_model, model_dir):
    """Returns a list of TrainHook objects for a run that writes to model_dir."""
