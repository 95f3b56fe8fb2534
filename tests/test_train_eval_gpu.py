"""GPU integration tests through the public T2R API, mirroring the reference's T2RModelFixture
(utils/t2r_test_fixture.py:42-140; research/qtopt/t2r_models_test.py:39-52): `random_train` =
DefaultRandomInputGenerator + train_eval_model for 2 steps at batch 2, `recordio_train` on TFRecords,
PREDICT with action_batch_size=64."""
import glob
import io
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(**kwargs):
  from tensor2robot_b200.research.qtopt import t2r_models
  return t2r_models.Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom(**kwargs)


def test_random_train_two_steps(tmp_path):
  from tensor2robot_b200.input_generators import default_input_generator as gens
  from tensor2robot_b200.utils import train_eval
  model = _model()
  out = train_eval.train_eval_model(t2r_model=model, input_generator_train=gens.DefaultRandomInputGenerator(batch_size=2),
                                    max_train_steps=2, model_dir=str(tmp_path))
  assert out['global_step'] == 2 and np.isfinite(out['loss'])
  files = sorted(os.path.basename(p) for p in glob.glob(str(tmp_path / 'model.ckpt-*.pt')))
  assert files == ['model.ckpt-0.pt', 'model.ckpt-2.pt']
  # every trainable variable moved; moving statistics were updated by the fused BN kernels
  state = torch.load(str(tmp_path / 'model.ckpt-2.pt'), weights_only=False)
  init = torch.load(str(tmp_path / 'model.ckpt-0.pt'), weights_only=False)
  name = 'Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom/conv2/weights'
  assert state['variables'][name].shape == (5, 5, 64, 64)                      # reference (HWIO) layout
  assert np.abs(state['variables'][name] - init['variables'][name]).max() > 0
  mm = 'Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom/conv2/BatchNorm/moving_mean'
  assert np.abs(state['variables'][mm]).max() > 0
  # resuming continues from the checkpoint instead of restarting
  model2 = _model()
  out2 = train_eval.train_eval_model(t2r_model=model2, input_generator_train=gens.DefaultRandomInputGenerator(batch_size=2),
                                     max_train_steps=3, model_dir=str(tmp_path))
  assert out2['global_step'] == 3


def test_predict_with_action_batch(tmp_path):
  from tensor2robot_b200.input_generators import default_input_generator as gens
  from tensor2robot_b200.utils import train_eval
  model = _model(action_batch_size=64)
  preds = train_eval.predict_from_model(t2r_model=model, input_generator_predict=gens.DefaultRandomInputGenerator(batch_size=2),
                                        model_dir=None)
  q = next(preds)['q_predicted']
  assert tuple(q.shape) == (2, 64)
  assert float(q.min()) >= 0 and float(q.max()) <= 1
  assert model.global_step == 0


def _write_replay(path, n, seed=0):
  from PIL import Image
  from oracle import tfrecord
  rng = np.random.RandomState(seed)
  records = []
  for i in range(n):
    img = rng.randint(0, 256, (64, 80, 3)).astype(np.uint8)
    img = np.asarray(Image.fromarray(img).resize((640, 512), Image.BILINEAR))   # low-pass content
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format='JPEG', quality=90)
    feats = {'image_1': buf.getvalue(), 'world_vector': rng.uniform(-1, 1, 3), 'vertical_rotation': rng.uniform(-1, 1, 2),
             'grasp_success': [float(rng.uniform() < 0.3)]}
    for k in ('close_gripper', 'open_gripper', 'terminate_episode', 'gripper_closed'):
      feats[k] = [float(rng.uniform() < 0.5)]
    feats['height_to_bottom'] = [float(rng.uniform())]
    records.append(tfrecord.make_example({k: (v if isinstance(v, bytes) else list(np.float32(v))) for k, v in feats.items()}))
  tfrecord.write_tfrecords(path, records)


def test_record_train_and_eval(tmp_path):
  from tensor2robot_b200.input_generators import default_input_generator as gens
  from tensor2robot_b200.utils import train_eval
  data = str(tmp_path / 'replay.tfrecord')
  _write_replay(data, 12)
  model = _model()
  out = train_eval.train_eval_model(
      t2r_model=model, input_generator_train=gens.DefaultRecordInputGenerator(file_patterns=data, batch_size=4, seed=0),
      input_generator_eval=gens.DefaultRecordInputGenerator(file_patterns=data, batch_size=4),
      max_train_steps=2, eval_steps=2, model_dir=str(tmp_path / 'run'))
  assert out['global_step'] == 2 and np.isfinite(out['loss'])
  assert out['eval']['steps'] == 2 and np.isfinite(out['eval']['loss'])


def test_record_train_with_device_jpeg_decoder(tmp_path):
  """The same TFRecord -> parse -> decode -> preprocess -> train path with the split JPEG decoder
  (Huffman on host, IDCT / colour on the GPU): the parser hands CUDA uint8 frames to the preprocessor.  The
  decoders are bit-identical (tests/test_jpeg.py); the first loss is only compared loosely because the TRAIN
  record stream is non-deterministic by design (utils/tfdata.py:683-685) and BN statistics use atomics."""
  from tensor2robot_b200.input_generators import default_input_generator as gens
  from tensor2robot_b200.utils import tfdata, train_eval
  data = str(tmp_path / 'replay.tfrecord')
  _write_replay(data, 8)
  losses = {}
  from tensor2robot_b200.preprocessors import image_transformations
  for kind in ('host', 'device'):
    tfdata.set_image_decoder(kind)
    image_transformations.seed(0)          # same random crop offsets in both runs
    try:
      out = train_eval.train_eval_model(
          t2r_model=_model(), input_generator_train=gens.DefaultRecordInputGenerator(file_patterns=data, batch_size=4,
                                                                                    seed=0),
          max_train_steps=1, model_dir=str(tmp_path / ('run_' + kind)))
    finally:
      tfdata.set_image_decoder('auto')
    assert out['global_step'] == 1 and np.isfinite(out['loss'])
    losses[kind] = out['loss']
  assert abs(losses['host'] - losses['device']) < 0.05 * abs(losses['host']), losses


def test_engine_requires_cuda_tensors():
  from tensor2robot_b200 import _lib, nn
  with pytest.raises(_lib.T2RError):
    nn.max_pool2d(torch.zeros((1, 4, 4, 8), dtype=torch.bfloat16), 2, 2)
