"""bin/run_t2r_trainer: flag parsing / model resolution on CPU, one short training run on the GPU."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, 'golden', 'pose_env_test_data.tfrecord')
MODEL = 'tensor2robot_b200.research.pose_env.pose_env_models:PoseEnvRegressionModel'


def test_flags_and_model_resolution():
  from tensor2robot_b200.bin import run_t2r_trainer as cli
  from tensor2robot_b200.input_generators import default_input_generator as gens
  from tensor2robot_b200.research.pose_env import pose_env_models as pm
  assert cli.resolve(MODEL) is pm.PoseEnvRegressionModel
  with pytest.raises(ValueError):
    cli.resolve('tensor2robot_b200.research.pose_env.pose_env_models')
  args = cli.build_parser().parse_args(['--model', MODEL, '--model_kwargs', json.dumps({'action_size': 2}),
                                        '--train_file_patterns', FIXTURE, '--batch_size', '4', '--max_train_steps', '7'])
  assert args.max_train_steps == 7 and args.eval_steps == 100 and args.image_decoder == 'auto'
  train, evaluation = cli.make_generators(args, (0, 1))
  assert isinstance(train, gens.DefaultRecordInputGenerator) and evaluation is None
  args = cli.build_parser().parse_args(['--model', MODEL])
  assert isinstance(cli.make_generators(args, (0, 1))[0], gens.DefaultRandomInputGenerator)


@pytest.mark.gpu
def test_cli_trains_on_the_fixture(tmp_path, capsys):
  from tensor2robot_b200.bin import run_t2r_trainer as cli
  result = cli.main(['--model', MODEL, '--train_file_patterns', FIXTURE, '--eval_file_patterns', FIXTURE,
                     '--batch_size', '4', '--max_train_steps', '2', '--eval_steps', '1', '--model_dir', str(tmp_path),
                     '--image_decoder', 'device'])
  from tensor2robot_b200.utils import tfdata
  tfdata.set_image_decoder('auto')
  assert result['global_step'] == 2 and result['eval']['steps'] == 1
  assert json.loads(capsys.readouterr().out.strip().splitlines()[-1])['global_step'] == 2
  assert os.path.exists(str(tmp_path / 'model.ckpt-2.pt'))
