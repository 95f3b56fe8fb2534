"""BC-Z pose assembly and losses (research/bcz/model.py:321-585) against an independent numpy restatement of the
TF-1.x semantics they rely on: tf.losses.huber_loss / log_loss with SUM_BY_NONZERO_WEIGHTS, quaternion
normalisation + Hamilton product ([x, y, z, w], tensorflow_graphics), residual composition, stop-token masking.
CPU only: these are host-scale torch functions."""
import numpy as np
import torch

from tensor2robot_b200.research.bcz import model as bcz
from tensor2robot_b200.utils import tensorspec_utils as utils


def _np_huber(label, pred, w):
  e = np.abs(pred - label)
  l = np.where(e <= 1.0, 0.5 * e * e, e - 0.5)
  w = np.broadcast_to(np.asarray(w, np.float64), l.shape)
  return (l * w).sum() / max((w != 0).sum(), 1)


def _np_logloss(label, pred, w):
  l = -label * np.log(pred + 1e-7) - (1 - label) * np.log(1 - pred + 1e-7)
  w = np.broadcast_to(np.asarray(w, np.float64), l.shape)
  return (l * w).sum() / max((w != 0).sum(), 1)


def test_infer_and_training_outputs_match_restatement():
  rng = np.random.RandomState(0)
  b, wpts = 5, 3
  comps = [('xyz', 3, True, 100.), ('quaternion', 4, True, 10.), ('target_close', 1, False, 1.)]
  net = {'xyz_residual': rng.standard_normal((b, wpts, 3)), 'quaternion_residual': rng.standard_normal((b, wpts, 4)),
         'target_close': rng.standard_normal((b, wpts, 1))}
  cur_q = rng.standard_normal((b, 4)); cur_q /= np.linalg.norm(cur_q, axis=-1, keepdims=True)
  present = {'xyz': rng.standard_normal((b, 3)), 'quaternion': cur_q}
  features = utils.TensorSpecStruct()
  features['present/xyz'] = torch.from_numpy(present['xyz'])
  features['present/quaternion'] = torch.from_numpy(present['quaternion'])
  out = bcz.infer_outputs(features, {k: torch.from_numpy(v) for k, v in net.items()}, comps, rescale_target_close=False)
  # xyz: residual added to the present pose
  np.testing.assert_allclose(out['action/xyz'].numpy(), net['xyz_residual'] + present['xyz'][:, None, :], rtol=1e-12)
  # quaternion: normalise, then present * predicted (Hamilton product, xyzw)
  qn = np.linalg.norm(net['quaternion_residual'], axis=-1, keepdims=True)
  q = net['quaternion_residual'] / qn
  a = np.broadcast_to(present['quaternion'][:, None, :], q.shape)
  x1, y1, z1, w1 = [a[..., i] for i in range(4)]
  x2, y2, z2, w2 = [q[..., i] for i in range(4)]
  want_q = np.stack([x1 * w2 + y1 * z2 - z1 * y2 + w1 * x2, -x1 * z2 + y1 * w2 + z1 * x2 + w1 * y2,
                     x1 * y2 - y1 * x2 + z1 * w2 + w1 * z2, -x1 * x2 - y1 * y2 - z1 * z2 + w1 * w2], -1)
  np.testing.assert_allclose(out['action/quaternion'].numpy(), want_q, rtol=1e-10)
  np.testing.assert_allclose(np.linalg.norm(out['action/quaternion'].numpy(), axis=-1), 1.0, rtol=1e-10)
  np.testing.assert_allclose(out['quaternion_norm'].numpy(), qn, rtol=1e-12)
  np.testing.assert_allclose(out['action/target_close'].numpy(), 1 / (1 + np.exp(-net['target_close'])), rtol=1e-10)
  assert tuple(out['action_trajectory'].shape) == (b, wpts, 8)
  np.testing.assert_allclose(bcz.xyz_action_trajectory(out).numpy(),
                             np.concatenate([out['action/xyz'].numpy(), want_q], -1), rtol=1e-10)

  # ---- losses on the UNMODIFIED head outputs (except the quaternion, overwritten by infer_outputs) ----
  labels = utils.TensorSpecStruct()
  lab = {'xyz_residual': rng.standard_normal((b, wpts, 3)), 'quaternion_residual': rng.standard_normal((b, wpts, 4)),
         'target_close': (rng.uniform(size=(b, wpts, 1)) < 0.5).astype(np.float64),
         'stop_token': (rng.uniform(size=(b, wpts, 1)) < 0.3).astype(np.float64)}
  for k, v in lab.items():
    labels['future/' + k] = torch.from_numpy(v)
  netd = {k: torch.from_numpy(v) for k, v in net.items()}
  netd['quaternion_norm'] = torch.from_numpy(qn)
  loss, train = bcz.training_outputs(labels, netd, comps, quaternion_penalty=0.01, loss_name='huber')
  mask = 1.0 - lab['stop_token']
  want = {
      'xyz_loss': _np_huber(lab['xyz_residual'], net['xyz_residual'], 100. * mask * np.ones_like(net['xyz_residual'])),
      'quaternion_loss': _np_huber(lab['quaternion_residual'], net['quaternion_residual'],
                                   10. * mask * np.ones_like(net['quaternion_residual'])),
      'target_close_loss': _np_logloss(lab['target_close'], 1 / (1 + np.exp(-net['target_close'])), 1. * mask),
      'quaternion_norm_loss': _np_huber(np.ones_like(qn), qn, 0.01 * mask),
  }
  for k, v in want.items():
    np.testing.assert_allclose(float(train[k]), v, rtol=1e-9, err_msg=k)
  np.testing.assert_allclose(float(loss), sum(want.values()), rtol=1e-9)
  np.testing.assert_allclose(float(train['first_xyz_error']),
                             _np_huber(lab['xyz_residual'][:, 0], net['xyz_residual'][:, 0], 100.), rtol=1e-9)
  # all steps stopped -> every weight is zero -> the weighted losses vanish (SUM_BY_NONZERO_WEIGHTS of nothing)
  labels['future/stop_token'] = torch.ones((b, wpts, 1), dtype=torch.float64)
  loss0, _ = bcz.training_outputs(labels, netd, comps)
  assert float(loss0) == 0.0
  # piecewise scaling only kicks in above 1
  f = bcz.piecewise_scaled_huber(lambda **kw: torch.tensor(kw['v']))
  assert float(f(v=0.5)) == 0.5 and abs(float(f(v=3.0)) - (0.2 + 2.8 * 0.001)) < 1e-6


def test_bcz_model_specs():
  """research/bcz/model.py:690-790: feature / label specs and what the preprocessor asks the parser for."""
  from tensor2robot_b200.research.bcz import model as bcz
  from tensor2robot_b200.utils import dtypes
  m = bcz.BCZModel(image_size=(100, 100), num_waypoints=10)
  f = m.get_feature_specification('train')
  assert f['image'].shape == (100, 100, 3) and f['image'].name == 'present/image/encoded'
  assert f['present/target_close'].name == 'present/sensed_close'
  assert f['subtask_id'].dtype == dtypes.int64 and 'sentence_embedding' not in f.keys()
  l = m.get_label_specification('train')
  assert sorted(l.keys()) == ['future/quaternion', 'future/target_close', 'future/xyz_residual']
  assert l['future/xyz_residual'].shape == (10, 3) and l['future/xyz_residual'].name == 'future/xyz_residual'
  pin = m.preprocessor.get_in_feature_specification('train')
  assert pin['image'].shape == (512, 640, 3) and pin['image'].dtype == dtypes.uint8
  assert 'original_image' not in pin.keys()
  lang = bcz.BCZModel(cond_modality=bcz.ConditionMode.LANGUAGE_EMBEDDING)
  assert lang.get_feature_specification('train')['sentence_embedding'].shape == (512,)
  assert bcz.MIN_GRIPPER_CLOSE == 0.2 and bcz.NUM_DEBUG_TASKS == 21


def test_tf_one_hot_semantics():
  import torch
  from tensor2robot_b200.research.bcz import model as bcz
  oh = bcz._one_hot(torch.tensor([0, 20, 21, 254]), 21)
  assert oh.shape == (4, 21) and oh.sum(1).tolist() == [1.0, 1.0, 0.0, 0.0]


def test_stop_state_loss_and_spec():
  """model.py:462-473, 566-573: softmax cross-entropy of the one-hot stop state, weighted per class."""
  import torch
  from tensor2robot_b200.research.bcz import model as bcz
  rng = np.random.RandomState(3)
  logits = rng.standard_normal((5, 3)).astype(np.float32)
  state = np.array([0, 2, 1, 1, 7])                       # 7 is out of range: tf.one_hot row of zeros, weight 0
  class_weights = [1.0, 0.5, 3.0]
  onehot = np.zeros((5, 3), np.float32)
  for i, s_ in enumerate(state):
    if s_ < 3:
      onehot[i, s_] = 1
  logp = logits - np.log(np.exp(logits).sum(1, keepdims=True))
  ce = -(onehot * logp).sum(1)
  w = (onehot * np.array(class_weights, np.float32)).sum(1)
  want = (ce * w).sum() / (w != 0).sum()
  got = bcz.compute_stop_state_loss(torch.from_numpy(onehot), torch.from_numpy(logits), class_weights)
  np.testing.assert_allclose(float(got), want, rtol=1e-6)
  np.testing.assert_array_equal(bcz._one_hot(torch.from_numpy(state), 3).numpy(), onehot)
  m = bcz.BCZModel(predict_stop=True, stop_state_class_weights=class_weights)
  spec = m.get_label_specification('train')['future/stop_state']
  assert spec.shape == () and spec.name == 'present/stop_state'
  assert 'future/stop_state' not in bcz.BCZModel().get_label_specification('train').keys()


def test_mixup_reverse_formula():
  """research/bcz/model.py:164-172: lmbda * x + (1 - lmbda) * tf.reverse(x, axis=[0])."""
  import torch
  from tensor2robot_b200.research.bcz import model as bcz
  x = torch.arange(24, dtype=torch.float32).reshape(4, 3, 2)
  y = bcz.mixup_reverse(x, 0.25)
  np.testing.assert_allclose(y.numpy(), 0.25 * x.numpy() + 0.75 * x.numpy()[::-1])
  assert bcz.BCZPreprocessor.__init__.__defaults__[4] == 0.0          # mixup_alpha is off by default
