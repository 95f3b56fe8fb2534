"""Host-side logic of the data-parallel path on CPU: 2 ranks over gloo (127.0.0.1).

Covers what does not need a GPU: rank sharding of the record stream (disjoint and
together complete), the single gradient all-reduce + 1/N scaling convention, and the initial
parameter broadcast.  The NCCL path itself is exercised by bench.py --gpus N on the B200 box."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, 'golden', 'pose_env_test_data.tfrecord')


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world, port, out_dir):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from tensor2robot_b200 import engine
  from tensor2robot_b200.input_generators import default_input_generator as gens
  from tensor2robot_b200.utils import dtypes
  from tensor2robot_b200.utils import tensorspec_utils as utils
  spec = utils.TensorSpecStruct(pose=utils.ExtendedTensorSpec((2,), dtypes.float32, 'pose'))
  labels = utils.TensorSpecStruct(reward=utils.ExtendedTensorSpec((1,), dtypes.float32, 'reward'))
  assert engine.shard_for_rank() == (rank, world)
  g = gens.DefaultRecordInputGenerator(file_patterns=FIXTURE, batch_size=10, shard=engine.shard_for_rank())
  g.set_feature_specifications(spec, spec)
  g.set_label_specifications(labels, labels)
  poses = np.concatenate([f.pose for f, _ in g.create_dataset('eval')])
  # a stand-in "gradient": per-replica sum over its local batch, written into a flat buffer
  flat_grad = torch.from_numpy(poses.sum(0)).clone()
  scale = engine.reduce_gradients(flat_grad)
  params = torch.full((4,), float(rank + 1))
  dist.broadcast(params, src=0)
  # the chief alone writes checkpoints / runs chief hooks (replicas hold identical parameters)
  from tensor2robot_b200.utils import train_eval

  class _Model(object):
    global_step = 7

    def state_dict(self):
      return {'variables': {}, 'global_step': 7, 'rank': rank}

  path = train_eval.save_checkpoint(_Model(), os.path.join(out_dir, 'ckpt'))
  dist.barrier()
  wrote = torch.load(path, weights_only=False)['rank'] if os.path.exists(path) else None
  np.save(os.path.join(out_dir, 'rank%d.npy' % rank), {'poses': poses, 'grad': flat_grad.numpy(), 'scale': scale,
                                                       'params': params.numpy(), 'ckpt_rank': wrote}, allow_pickle=True)
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_sharding_and_gradient_reduction(tmp_path):
  world, port = 2, _free_port()
  mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
  golden = np.load(os.path.join(HERE, 'golden', 'pose_env_golden.npz'))['pose']
  res = [np.load(os.path.join(str(tmp_path), 'rank%d.npy' % r), allow_pickle=True).item() for r in range(world)]
  np.testing.assert_array_equal(res[0]['poses'], golden[0::2])          # disjoint ...
  np.testing.assert_array_equal(res[1]['poses'], golden[1::2])          # ... and complete
  total = golden.astype(np.float32).reshape(-1, 2)
  for r in res:
    assert r['scale'] == 0.5
    np.testing.assert_allclose(r['grad'], total.sum(0), rtol=1e-5)      # identical sum on every rank
    np.testing.assert_allclose(r['grad'] * r['scale'], total.sum(0) / 2, rtol=1e-5)
    np.testing.assert_array_equal(r['params'], np.full(4, 1.0, np.float32))   # rank 0's initial values
    assert r['ckpt_rank'] == 0                                          # one checkpoint file, written by rank 0


def test_single_process_is_identity():
  from tensor2robot_b200 import engine
  g = torch.ones(8)
  assert engine.reduce_gradients(g) == 1.0 and engine.shard_for_rank() == (0, 1)
  assert torch.equal(g, torch.ones(8))
