"""Trainer entry point (bin/run_t2r_trainer.py:20-37 of the reference: parse the configuration, call
train_eval.train_eval_model).  gin is not available, so the configuration is command-line flags: the model as
`package.module:ClassName` plus JSON keyword arguments, the record files, and the train_eval_model arguments.

  python -m tensor2robot_b200.bin.run_t2r_trainer \\
      --model tensor2robot_b200.research.pose_env.pose_env_models:PoseEnvRegressionModel \\
      --train_file_patterns tests/golden/pose_env_test_data.tfrecord --batch_size 32 --max_train_steps 100 \\
      --model_dir /tmp/pose_env

Under torchrun (one process per GPU) it joins the NCCL group, takes cuda:LOCAL_RANK and reads its shard of the
record files; the gradient all-reduce is the model's own train_step."""
import argparse
import importlib
import json
import logging
import os
import sys


def resolve(path):
  """'package.module:Attribute' -> the attribute."""
  module, _, attr = path.partition(':')
  if not attr:
    raise ValueError('expected package.module:ClassName, got %r' % path)
  return getattr(importlib.import_module(module), attr)


def build_parser():
  p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
  p.add_argument('--model', required=True, help='T2R model class as package.module:ClassName')
  p.add_argument('--model_kwargs', default='{}', help='JSON keyword arguments of the model constructor')
  p.add_argument('--train_file_patterns', default=None, help='comma separated TFRecord patterns; random inputs if omitted')
  p.add_argument('--eval_file_patterns', default=None)
  p.add_argument('--batch_size', type=int, default=32, help='per-process (per-GPU) batch size')
  p.add_argument('--max_train_steps', type=int, default=1000)
  p.add_argument('--eval_steps', type=int, default=100)
  p.add_argument('--model_dir', default='/tmp/t2r_b200')
  p.add_argument('--log_every_n_steps', type=int, default=100)
  p.add_argument('--image_decoder', choices=('auto', 'host', 'device'), default='auto',
                 help="JPEG decoder of the record parser ('auto': the split host/GPU decoder when a GPU is present)")
  return p


def make_generators(args, shard):
  from tensor2robot_b200.input_generators import default_input_generator as gens
  if args.train_file_patterns:
    train = gens.DefaultRecordInputGenerator(file_patterns=args.train_file_patterns, batch_size=args.batch_size, shard=shard)
  else:
    train = gens.DefaultRandomInputGenerator(batch_size=args.batch_size)
  evaluation = None
  if args.eval_file_patterns:
    evaluation = gens.DefaultRecordInputGenerator(file_patterns=args.eval_file_patterns, batch_size=args.batch_size,
                                                  shard=shard)
  return train, evaluation


def main(argv=None):
  args = build_parser().parse_args(argv)
  logging.basicConfig(level=logging.INFO)
  import torch
  import torch.distributed as dist
  from tensor2robot_b200 import engine
  from tensor2robot_b200.utils import tfdata
  from tensor2robot_b200.utils import train_eval
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local_rank)
  if world > 1 and not dist.is_initialized():
    dist.init_process_group('nccl')
  tfdata.set_image_decoder(args.image_decoder)
  model = resolve(args.model)(**json.loads(args.model_kwargs))
  train, evaluation = make_generators(args, engine.shard_for_rank())
  result = train_eval.train_eval_model(t2r_model=model, input_generator_train=train, input_generator_eval=evaluation,
                                       max_train_steps=args.max_train_steps, eval_steps=args.eval_steps,
                                       model_dir=args.model_dir, log_every_n_steps=args.log_every_n_steps)
  if int(os.environ.get('RANK', '0')) == 0:
    print(json.dumps(result))
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  return result


if __name__ == '__main__':
  main(sys.argv[1:])
