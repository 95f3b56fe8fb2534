"""BC-Z through the public T2R API on the GPU (research/bcz/model_test.py: random_train of BCZModel)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(**kwargs):
  from tensor2robot_b200.research.bcz import model as bcz
  pre = lambda **kw: bcz.BCZPreprocessor(image_size=(96, 96), crop_size=(104, 128), input_size=(112, 144),
                                         rescale_gripper=True, **kw)
  return bcz.BCZModel(image_size=(96, 96), input_size=(112, 144), resnet_size=18, num_waypoints=3,
                      preprocessor_cls=pre, **kwargs)


@pytest.mark.parametrize('cond', ['onehot', 'language', 'onehot+stop', 'spatial'])
def test_bcz_random_train(tmp_path, cond):
  from tensor2robot_b200.input_generators import default_input_generator as gens
  from tensor2robot_b200.research.bcz import model as bcz
  from tensor2robot_b200.utils import train_eval
  mode = bcz.ConditionMode.LANGUAGE_EMBEDDING if cond == 'language' else bcz.ConditionMode.ONEHOT_TASKID
  extra = dict(predict_stop=True, stop_state_class_weights=[1.0, 2.0, 2.0]) if cond.endswith('stop') else {}
  if cond == 'spatial':
    extra = dict(network_fn=bcz.spatial_softmax_network)
  model = _model(cond_modality=mode, **extra)
  out = train_eval.train_eval_model(t2r_model=model, input_generator_train=gens.DefaultRandomInputGenerator(batch_size=4),
                                    max_train_steps=2, model_dir=str(tmp_path))
  assert out['global_step'] == 2 and np.isfinite(out['loss'])
  state = torch.load(str(tmp_path / 'model.ckpt-2.pt'), weights_only=False)
  init = torch.load(str(tmp_path / 'model.ckpt-0.pt'), weights_only=False)
  if cond == 'spatial':
    assert init['variables']['vision_model/pose_fc0/weights'].shape == (64 + 21 + 10, 100)     # points + task + bias transform
    assert init['variables']['vision_model/pose_fc2/weights'].shape == (100, (3 + 4 + 1) * 3)
    assert all(np.abs(state['variables'][k] - init['variables'][k]).max() > 0 for k in init['variables'])
    return
  moved = [k for k in init['variables'] if 'film' in k and np.abs(state['variables'][k] - init['variables'][k]).max() > 0]
  assert moved, 'the FiLM generator must receive gradients'
  if cond.endswith('stop'):      # the stop-state head exists under the reference's scope names
    assert 'predict_stop/Stack/fully_connected_1/LayerNorm/gamma' in init['variables']
    assert init['variables']['predict_stop/fully_connected/weights'].shape == (100, 3)


def test_bcz_predict(tmp_path):
  from tensor2robot_b200.input_generators import default_input_generator as gens
  from tensor2robot_b200.utils import train_eval
  model = _model()
  preds = next(train_eval.predict_from_model(t2r_model=model,
                                             input_generator_predict=gens.DefaultRandomInputGenerator(batch_size=2),
                                             model_dir=None))
  assert tuple(preds['action/xyz'].shape) == (2, 3, 3) and tuple(preds['action/quaternion'].shape) == (2, 3, 4)
  q = preds['action/quaternion'].float()
  assert tuple(preds['action_trajectory'].shape) == (2, 3, 8)
  assert torch.allclose(q.norm(dim=-1), torch.ones(2, 3, device=q.device), atol=1e-4)   # unit quaternions
  g = preds['action/target_close'].float()
  assert float(g.min()) >= 0.2 - 1e-6 and float(g.max()) <= 1 + 1e-6                     # rescaled gripper range
