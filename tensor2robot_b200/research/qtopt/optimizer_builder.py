"""BuildOpt: the QT-Opt optimizer from hyper-parameters (research/qtopt/optimizer_builder.py:25-96)."""
from tensor2robot_b200.models import optimizers


class HParams(dict):
  """Minimal stand-in for contrib_training.HParams: attribute access + .get."""

  def __getattr__(self, name):
    try:
      return self[name]
    except KeyError:
      raise AttributeError(name)

  def __setattr__(self, name, value):
    self[name] = value


def BuildOpt(hparams):  # pylint: disable=invalid-name
  """Staircase exponential LR decay every int(examples_per_epoch / batch_size * num_epochs_per_decay)
  steps; momentum (default) / rmsprop(decay, momentum, epsilon) / adam(beta1=momentum); optionally wrapped in
  MovingAverageOptimizer."""
  decay_steps = int(hparams.examples_per_epoch / hparams.batch_size * hparams.num_epochs_per_decay)
  learning_rate = optimizers.create_exp_decaying_learning_rate(
      hparams.learning_rate, decay_steps, hparams.learning_rate_decay_factor, staircase=True)
  optimizer = hparams.optimizer
  if optimizer == 'momentum':
    opt = optimizers.MomentumOptimizer(learning_rate, hparams.momentum)
  elif optimizer == 'rmsprop':
    opt = optimizers.RMSPropOptimizer(learning_rate, decay=hparams.rmsprop_decay, momentum=hparams.momentum,
                                      epsilon=hparams.rmsprop_epsilon)
  else:
    opt = optimizers.AdamOptimizer(learning_rate, beta1=hparams.momentum, beta2=hparams.get('adam_beta2', 0.999),
                                   epsilon=hparams.get('adam_epsilon', 1e-8))
  if hparams.use_avg_model_params:
    return optimizers.MovingAverageOptimizer(opt, average_decay=hparams.model_weights_averaging)
  return opt
