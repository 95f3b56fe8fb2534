"""MockT2RModel / MockInputGenerator (utils/mocks.py) and the cross-framework pin they allow: the reference's fixture
checkpoint - written by TensorFlow after 1100 training steps of this model on this dataset - is read by
utils/tf_checkpoint.py, and must separate the dataset (a) through the numpy oracle and (b) through the CUDA kernels
after being loaded into MockT2RModel by variable name."""
import os

import numpy as np
import pytest

from oracle import mocks as oracle
from tensor2robot_b200.utils import mocks
from tensor2robot_b200.utils import tf_checkpoint

HERE = os.path.dirname(os.path.abspath(__file__))
PREFIX = os.path.join(HERE, 'golden', 'mock_savedmodel_variables', 'variables')


def _tf_weights():
  reader = tf_checkpoint.load_checkpoint(PREFIX)
  return {n: reader.get_tensor(n) for n in reader.get_variable_to_shape_map()}


def test_tensorflow_trained_weights_separate_the_mock_dataset():
  features, labels = mocks.MockInputGenerator(batch_size=32).create_numpy_data()
  assert features.shape == (128, 3) and labels.sum() == 64 and (features[:64] > 0.2 - 1e-9).all() and (features[64:] < -0.2 + 1e-9).all()
  logits = oracle.forward(_tf_weights(), features)
  assert ((logits > 0) == (labels > 0.5)).all()
  assert oracle.categorical_hinge(labels, logits) < 1e-2         # trained to the hinge margin: |logit| >= 1 nearly everywhere
  assert (np.abs(logits) > 0.9).mean() > 0.95


def test_oracle_matches_the_inference_graph_tensorflow_exported():
  """The strongest pin the reference's fixtures allow without TensorFlow: its exported SavedModel
  (test_data/mock_exported_savedmodel/saved_model.pb, written by TF from utils/mocks.py:150-176) is evaluated op by op
  with numpy (tests/graphdef_interp.py) on the checkpoint beside it, and the restatement in oracle/mocks.py must
  reproduce it: same op order (dense -> bias -> elu -> moving-statistics batch norm with epsilon 1e-3, three times,
  then the linear head), same numbers."""
  import graphdef_interp
  graph = graphdef_interp.Graph(open(os.path.join(HERE, 'golden', 'mock_saved_model.pb'), 'rb').read())
  out = 'MockT2RModel.dense.4/BiasAdd'
  block = ['MatMul', 'BiasAdd', 'Elu', 'AddV2', 'Rsqrt', 'Mul', 'Mul', 'Mul', 'Sub', 'AddV2']
  assert graph.ops_on_path(out) == block * 3 + ['MatMul', 'BiasAdd']
  weights = _tf_weights()
  rng = np.random.RandomState(0)
  x = np.concatenate([mocks.MockInputGenerator(batch_size=32).create_numpy_data()[0],
                      rng.uniform(-2, 2, (64, 3))]).astype(np.float32)
  want = graph.run(out, {'measured_position': x}, {k: np.asarray(v, np.float32) for k, v in weights.items()})
  eps = graph.run('MockT2RModel.batch_norm.0/batchnorm/add/y', {}, {})
  assert abs(float(eps) - 1e-3) < 1e-9
  got = oracle.forward(weights, x)
  assert want.shape == got.shape == (x.shape[0], 1)
  np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6)       # fp32 graph evaluation against the float64 restatement


def test_input_generator_batches():
  from tensor2robot_b200.utils import train_eval
  model = mocks.MockT2RModel()
  gen = mocks.MockInputGenerator(batch_size=48)
  train_eval.provide_input_generator_with_model_information(gen, model, 'eval')
  batches = list(gen.create_dataset('eval'))
  assert len(batches) == 2 and batches[0][0]['x'].shape == (48, 3) and batches[0][1]['y'].shape == (48, 1)   # remainder dropped
  assert batches[0][1]['y'].all() and batches[0][0]['x'].dtype == np.float32
  train = gen.create_dataset('train')
  seen = [next(train)[1]['y'].mean() for _ in range(6)]          # repeats and shuffles
  assert 0 < np.mean(seen) < 1
  multi = mocks.MockT2RModel(multi_dataset=True)
  assert sorted(multi.get_feature_specification('train').keys()) == ['x1', 'x2']
  assert multi.get_feature_specification('train').x2.dataset_key == 'dataset2'


@pytest.mark.gpu
def test_mock_model_with_tensorflow_weights_and_training(tmp_path):
  import torch
  from tensor2robot_b200.models import abstract_model
  from tensor2robot_b200.predictors import checkpoint_predictor
  from tensor2robot_b200.utils import train_eval
  features, labels = mocks.MockInputGenerator(batch_size=32).create_numpy_data()
  weights = _tf_weights()
  # (b) TensorFlow's weights, loaded by name, through the CUDA kernels
  model = mocks.MockT2RModel(init_from_checkpoint_fn=abstract_model.default_init_from_checkpoint_fn(PREFIX))
  predictor = checkpoint_predictor.CheckpointPredictor(t2r_model=model)
  predictor.init_randomly()
  names = sorted(model.variable_store.export_tf())
  assert names == sorted(n for n in weights if n != 'global_step')          # identical variable names
  logits = predictor.predict({'x': features.astype(np.float32)})['logit']
  want = oracle.forward(weights, features.astype(np.float32))
  assert np.abs(logits - want).max() < 1e-5
  assert ((logits > 0) == (labels > 0.5)).all()
  # gradients of the hinge loss against the oracle by central differences on two parameters
  vs = model.variable_store
  x = torch.from_numpy(features.astype(np.float32)).cuda()
  y = torch.from_numpy(labels.astype(np.float32)).cuda()
  shift = {k: (v * 0.7 if 'kernel' in k else v) for k, v in weights.items() if k != 'global_step'}   # off the margin
  vs.import_tf(shift)
  from tensor2robot_b200 import nn
  from tensor2robot_b200.utils import tensorspec_utils as tu
  with nn.variable_store(vs):
    out = model.inference_network_fn(tu.TensorSpecStruct(x=x), None, 'train')
    loss = model.model_train_fn(None, tu.TensorSpecStruct(y=y), out, 'train')
    vs.zero_grad()
    loss.backward()
  grads = vs.export_tf_grads()
  np.testing.assert_allclose(float(loss), oracle.categorical_hinge(labels, oracle.forward(shift, features.astype(np.float32))), rtol=1e-5)
  for name, index in (('MockT2RModel.dense.1/kernel', (3, 5)), ('MockT2RModel.batch_norm.0/gamma', (7,)),
                      ('MockT2RModel.dense.4/bias', (0,)), ('MockT2RModel.batch_norm.2/beta', (2,))):
    plus = {k: np.array(v, np.float64) for k, v in shift.items()}
    minus = {k: np.array(v, np.float64) for k, v in shift.items()}
    h = 1e-4
    plus[name][index] += h
    minus[name][index] -= h
    fd = (oracle.categorical_hinge(labels, oracle.forward(plus, features.astype(np.float32))) -
          oracle.categorical_hinge(labels, oracle.forward(minus, features.astype(np.float32)))) / (2 * h)
    assert abs(grads[name][index] - fd) < 2e-4 + 1e-3 * abs(fd), (name, grads[name][index], fd)
  # training from scratch separates the data (train_eval_test's mock pipeline)
  from tensor2robot_b200.models import optimizers
  fresh = mocks.MockT2RModel(create_optimizer_fn=lambda use_summaries: optimizers.AdamOptimizer(0.01))
  result = train_eval.train_eval_model(t2r_model=fresh, input_generator_train=mocks.MockInputGenerator(batch_size=32),
                                       input_generator_eval=mocks.MockInputGenerator(batch_size=32), max_train_steps=400,
                                       eval_steps=4, model_dir=str(tmp_path))
  assert result['global_step'] == 400 and result['eval']['steps'] == 4
  trained = checkpoint_predictor.CheckpointPredictor(t2r_model=fresh, checkpoint_dir=str(tmp_path))
  assert trained.restore()
  acc = ((trained.predict({'x': features.astype(np.float32)})['logit'] > 0) == (labels > 0.5)).mean()
  assert acc > 0.95, acc
