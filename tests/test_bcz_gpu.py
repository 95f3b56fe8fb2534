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


@pytest.mark.parametrize('loss_name', ['huber', 'mse'])
@pytest.mark.parametrize('with_stop', [False, True])
def test_training_outputs_fused_kernel_matches_eager(loss_name, with_stop):
  """research/bcz/model.py:476-585: every loss term, diagnostic and gradient of the one-launch kernel
  (t2r_weighted_losses) against the op-by-op restatement of the same formulas on CPU tensors."""
  from tensor2robot_b200.research.bcz import model as bcz
  from tensor2robot_b200.research.bcz import pose_components_lib
  from tensor2robot_b200.utils import tensorspec_utils
  rng = np.random.RandomState(11)
  b, w = 6, 5
  components = list(pose_components_lib.DEFAULT_ACTION_COMPONENTS)
  net_cpu, future = {}, tensorspec_utils.TensorSpecStruct()
  for name, size, residual, _ in components:
    key = name + '_residual' if residual else name
    net_cpu[key] = torch.from_numpy(rng.standard_normal((b, w, size)).astype(np.float32) * 1.5).requires_grad_(True)
    future[key] = torch.from_numpy(rng.uniform(0, 1, (b, w, size)).astype(np.float32))
  net_cpu['quaternion_norm'] = torch.from_numpy(rng.uniform(0.5, 2.5, (b, w, 1)).astype(np.float32)).requires_grad_(True)
  if with_stop:
    future['stop_token'] = torch.from_numpy((rng.uniform(size=(b, w, 1)) < 0.4).astype(np.float32))
  labels_cpu = tensorspec_utils.TensorSpecStruct(future=future)
  loss_cpu, out_cpu = bcz.training_outputs(labels_cpu, net_cpu, components, loss_name=loss_name)
  loss_cpu.backward()

  net_gpu = {k: v.detach().cuda().requires_grad_(True) for k, v in net_cpu.items()}
  labels_gpu = tensorspec_utils.TensorSpecStruct(
      future=tensorspec_utils.TensorSpecStruct([(k, v.cuda()) for k, v in future.items()]))
  from tensor2robot_b200 import _lib
  launches0 = _lib.launch_count()
  loss_gpu, out_gpu = bcz.training_outputs(labels_gpu, net_gpu, components, loss_name=loss_name)
  assert _lib.launch_count() - launches0 == 1                    # the whole loss tail is one kernel
  loss_gpu.backward()
  assert sorted(out_gpu) == sorted(out_cpu)
  np.testing.assert_allclose(float(loss_gpu), float(loss_cpu), rtol=2e-6)
  for k in out_cpu:
    np.testing.assert_allclose(out_gpu[k].detach().cpu().numpy(), out_cpu[k].detach().numpy(), rtol=1e-5, atol=1e-7,
                               err_msg=k)
  for k in net_cpu:
    np.testing.assert_allclose(net_gpu[k].grad.cpu().numpy(), net_cpu[k].grad.numpy(), rtol=1e-5, atol=1e-8, err_msg=k)


def test_mixup_preprocessing_blends_images_and_labels():
  """BCZPreprocessor(mixup_alpha > 0) in TRAIN mode (research/bcz/model.py:164-172): image and every future label are
  blended with the reversed batch by ONE Beta(alpha, alpha) draw; EVAL mode is untouched."""
  from tensor2robot_b200.research.bcz import model as bcz
  from tensor2robot_b200.utils import tensorspec_utils
  rng = np.random.RandomState(2)
  x = torch.from_numpy(rng.standard_normal((6, 5, 7)).astype(np.float32)).cuda()
  for lmbda in (0.0, 0.3, 1.0):
    got = bcz.mixup_reverse(x, lmbda).cpu().numpy()
    np.testing.assert_allclose(got, lmbda * x.cpu().numpy() + (1 - lmbda) * x.cpu().numpy()[::-1], rtol=1e-6, atol=1e-7)
  pre = lambda alpha, **kw: bcz.BCZPreprocessor(image_size=(96, 96), crop_size=(104, 128), input_size=(112, 144),
                                                binarize_gripper=False, mixup_alpha=alpha, **kw)
  outs = {}
  for alpha in (0.0, 0.4):
    model = bcz.BCZModel(image_size=(96, 96), input_size=(112, 144), resnet_size=18, num_waypoints=2,
                         preprocessor_cls=lambda a=alpha, **kw: pre(a, **kw))
    p = model.preprocessor
    rs = np.random.RandomState(5)
    feats = tensorspec_utils.make_random_numpy(p.get_in_feature_specification('train'), 4)
    labels = tensorspec_utils.make_random_numpy(p.get_in_label_specification('train'), 4)
    del rs
    to_dev = lambda st: tensorspec_utils.TensorSpecStruct(
        [(k, torch.from_numpy(np.ascontiguousarray(v)).cuda()) for k, v in tensorspec_utils.flatten_spec_structure(st).items()])
    np.random.seed(0)
    bcz._RNG = np.random.RandomState(9)
    from tensor2robot_b200.preprocessors import image_transformations
    image_transformations.seed(0)
    f, l = p.preprocess(to_dev(feats), to_dev(labels), 'train')
    outs[alpha] = (feats, labels, f, l)
  # identical inputs are not guaranteed across the two random draws, so check the blend relation within the mixup run
  feats, labels, f, l = outs[0.4]
  lab_in = np.asarray(labels['future/xyz_residual'])
  lab_out = l.future['xyz_residual'].cpu().numpy()
  # solve lambda from one element, then every element must satisfy the same blend
  num = lab_out - lab_in[::-1]
  den = lab_in - lab_in[::-1]
  lam = float(np.median(num[np.abs(den) > 1e-3] / den[np.abs(den) > 1e-3]))
  assert 0.0 <= lam <= 1.0
  np.testing.assert_allclose(lab_out, lam * lab_in + (1 - lam) * lab_in[::-1], rtol=1e-4, atol=1e-5)
  assert f.image.dtype == torch.bfloat16 and tuple(f.image.shape) == (4, 96, 96, 3)


def test_bcz_eval_metrics(tmp_path):
  """model_eval_fn (research/bcz/model.py:894-929) through train_eval_model's evaluation pass: streaming means of the
  train outputs, stop-state accuracy, gripper closing / opening classification metrics."""
  from tensor2robot_b200.input_generators import default_input_generator as gens
  from tensor2robot_b200.utils import train_eval
  model = _model(predict_stop=True, stop_state_class_weights=[1.0, 2.0, 2.0])
  out = train_eval.train_eval_model(t2r_model=model, input_generator_train=gens.DefaultRandomInputGenerator(batch_size=4),
                                    input_generator_eval=gens.DefaultRandomInputGenerator(batch_size=4),
                                    max_train_steps=1, eval_steps=3, model_dir=str(tmp_path))
  ev = out['eval']
  assert ev['steps'] == 3 and np.isfinite(ev['loss'])
  for name in ('closing', 'opening'):
    for kind in ('accuracy', 'auc', 'precision', 'recall', 'pos_freq'):
      assert 0.0 <= ev['%s_%s' % (name, kind)] <= 1.0 + 1e-9, (name, kind, ev)
  assert 0.0 <= ev['accuracy_stop_state'] <= 1.0
  means = [k for k in ev if k.startswith('mean_')]
  assert 'mean_first_xyz_error' in means and len(means) >= 4
  assert all(np.isfinite(ev[k]) for k in means)
