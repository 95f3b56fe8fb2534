"""Robot-side action selection on top of a predictor (contract of the reference's
policies/policies.py:30-184: `SelectAction(state, context, timestep)`, `reset`, `restore`, `init_randomly`,
`global_step`, `model_path`, `sample_action`).

Two policies are on the path this repository accelerates:

* `CEMPolicy` - arg-max of a Q critic over continuous actions by the cross-entropy method.  Host mode drives
  `predictor.predict` once per CEM iteration with `cem_samples` candidate actions (one kernel pass of
  [1 x samples] through the critic).  Device mode (`device_maximizer=`, see `engine.device_cem_selector`) hands the
  whole search to `engine.CEMTargetComputer`: the state tower runs once, its features stay staged in HBM and
  sampling / Q evaluation / elite refit are kernels - the same code that computes Bellman targets, at batch 1.
* `RegressionPolicy` - the action is the regression model's `inference_output`.

CEM semantics (pinned by goldens generated from the reference's utils/cross_entropy.py, tests/golden): mean 0 /
stddev 1 start, ascending stable ordering with the last `num_elites` refitting mean and np.std(ddof=1), and the
answer is the best of the LAST iteration's samples."""
import numpy as np

from tensor2robot_b200.utils import cross_entropy


class Policy(object):
  """Holds the predictor and forwards the bookkeeping calls to it; subclasses implement SelectAction."""

  def __init__(self, predictor=None):
    self._predictor = predictor

  def SelectAction(self, state, context, timestep):  # pylint: disable=invalid-name
    raise NotImplementedError('%s does not implement SelectAction' % type(self).__name__)

  def reset(self):
    """Start of an episode: stateless policies have nothing to do."""

  def _forward(self, method):
    if self._predictor is not None:
      getattr(self._predictor, method)()

  def init_randomly(self):
    self._forward('init_randomly')

  def restore(self):
    self._forward('restore')

  @property
  def model_path(self):
    return 'No model path defined.' if self._predictor is None else self._predictor.model_path

  @property
  def global_step(self):
    return 0 if self._predictor is None else self._predictor.global_step

  def sample_action(self, obs, explore_prob):
    """dql_grasping's run_env protocol: (action, debug); exploration is the caller's business here."""
    del explore_prob
    return self.SelectAction(obs, None, None), None


class CEMPolicy(Policy):
  """Cross-entropy-method arg-max over a critic's continuous action space."""

  def __init__(self, t2r_model, action_size=2, cem_iters=3, cem_samples=64, num_elites=10, pack_fn=None,
               device_maximizer=None, **parent_kwargs):
    """pack_fn(t2r_model, state, context, timestep, samples) -> numpy feature struct for predictor.predict
    (default: t2r_model.pack_features).  device_maximizer(state) -> (action [D], q) replaces the host loop."""
    super(CEMPolicy, self).__init__(**parent_kwargs)
    self._t2r_model = t2r_model
    self._population = (int(cem_samples), int(action_size))
    self._cem_iters, self._num_elites = int(cem_iters), int(num_elites)
    self._device_maximizer = device_maximizer
    self.sample_fn = self._gaussian_population
    self.pack_fn = pack_fn if pack_fn is not None else (
        lambda model, state, context, timestep, samples: model.pack_features(state, context, timestep, samples))

  def _gaussian_population(self, mean, stddev):
    return mean + stddev * np.random.standard_normal(self._population)

  @staticmethod
  def _refit(params, elites):
    del params
    elites = np.asarray(elites)
    return dict(mean=elites.mean(axis=0), stddev=elites.std(axis=0, ddof=1))

  def get_cem_action(self, objective_fn):
    """objective_fn(samples [S, D]) -> values [S].  Returns (arg-max sample of the last iteration, debug dict with
    'q_predicted', 'final_params' {'mean', 'stddev'} and 'best_idx')."""
    dims = self._population[1]
    samples, values, params = cross_entropy.CrossEntropyMethod(
        self.sample_fn, objective_fn, self._refit, dict(mean=np.zeros(dims), stddev=np.ones(dims)),
        num_elites=self._num_elites, num_iterations=self._cem_iters)
    best = int(np.argmax(values))
    return samples[best], dict(q_predicted=values[best], final_params=params, best_idx=best)

  def SelectAction(self, state, context, timestep):  # pylint: disable=invalid-name
    if self._device_maximizer is not None:
      action, _ = self._device_maximizer(state)
      return np.asarray(action)

    def q_values(samples):
      features = self.pack_fn(self._t2r_model, state, context, timestep, samples)
      return self._predictor.predict(features)['q_predicted']

    return self.get_cem_action(q_values)[0]


class RegressionPolicy(Policy):
  """The first row of the regression model's `inference_output` is the action."""

  def __init__(self, t2r_model, **parent_kwargs):
    super(RegressionPolicy, self).__init__(**parent_kwargs)
    self._t2r_model = t2r_model

  def SelectAction(self, state, context, timestep):  # pylint: disable=invalid-name
    features = self._t2r_model.pack_features(state, context, timestep)
    return self._predictor.predict(features)['inference_output'][0]
