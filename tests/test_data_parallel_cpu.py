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


class _FakeVar(object):

  def __init__(self, name, offset, numel, regularize):
    self.name, self.offset, self.numel, self.regularize, self.trainable = name, offset, numel, regularize, True


class _FakeStore(object):
  """What GradientReducer needs of nn.VariableStore: flat layout [regularised | rest], a listener slot."""

  def __init__(self):
    import collections
    sizes = [('w%d' % i, 64 * (i + 1), True) for i in range(6)] + [('b%d' % i, 64, False) for i in range(3)]
    self.vars, off = collections.OrderedDict(), 0
    for name, n, reg in sizes:
      self.vars[name] = _FakeVar(name, off, n, reg)
      off += n
    self.flat_grad = torch.zeros(off)
    self.finalized = True
    self.grad_listener = None

  def grad_ready(self, var):
    if self.grad_listener is not None:
      self.grad_listener(var)


def _reducer_worker(rank, world, port, out_dir):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from tensor2robot_b200 import engine
  vs = _FakeStore()
  reducer = engine.GradientReducer(vs, n_buckets=3)
  out = {'buckets': [(b[0], b[1], sorted(b[2])) for b in reducer.buckets]}
  for trial, skip in (('all', ()), ('partial', ('w0', 'b1'))):
    vs.flat_grad.copy_(torch.arange(vs.flat_grad.numel(), dtype=torch.float32) * (rank + 1))
    reducer.begin()
    for name in reversed(list(vs.vars)):          # backward order: last layer first
      if name not in skip:
        vs.grad_ready(vs.vars[name])
    early = reducer.launched_early
    scale = reducer.finish()
    out[trial] = {'grad': vs.flat_grad.clone().numpy(), 'scale': scale, 'early': early,
                  'listener_cleared': vs.grad_listener is None}
  np.save(os.path.join(out_dir, 'reducer%d.npy' % rank), out, allow_pickle=True)
  dist.barrier()
  dist.destroy_process_group()


def test_bucketed_gradient_reducer_two_ranks(tmp_path):
  """engine.GradientReducer: buckets cover the flat buffer exactly once, fire as soon as their variables are done
  (backward order), anything never marked is reduced by finish(); the sum is the same on both ranks."""
  world, port = 2, _free_port()
  mp.spawn(_reducer_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
  res = [np.load(os.path.join(str(tmp_path), 'reducer%d.npy' % r), allow_pickle=True).item() for r in range(world)]
  buckets = res[0]['buckets']
  assert len(buckets) == 3                                              # weights cut by size + the remainder
  assert [b[2] for b in buckets[:2]] == [['w0', 'w1', 'w2', 'w3'], ['w4', 'w5']]
  assert buckets[0][0] == 0 and all(buckets[i][1] == buckets[i + 1][0] for i in range(2))
  n = 64 * 21 + 3 * 64
  assert buckets[-1][1] == n and buckets[-1][2] == ['b0', 'b1', 'b2']
  want = np.arange(n, dtype=np.float32) * 3                             # rank 0 (x1) + rank 1 (x2)
  for r in res:
    for trial, early in (('all', 3), ('partial', 1)):
      np.testing.assert_array_equal(r[trial]['grad'], want)
      assert r[trial]['scale'] == 0.5 and r[trial]['early'] == early and r[trial]['listener_cleared']


def test_single_process_is_identity():
  from tensor2robot_b200 import engine
  g = torch.ones(8)
  assert engine.reduce_gradients(g) == 1.0 and engine.shard_for_rank() == (0, 1)
  assert torch.equal(g, torch.ones(8))


def _gather_worker(rank, world, port, out_dir):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from oracle import grasp2vec as oracle
  from tensor2robot_b200.research.grasp2vec import losses
  rng = np.random.RandomState(0)
  a_all = rng.standard_normal((world * 3, 8)) * 0.5
  p_all = rng.standard_normal((world * 3, 8)) * 0.5
  a = torch.from_numpy(a_all[rank * 3:(rank + 1) * 3]).requires_grad_(True)
  p = torch.from_numpy(p_all[rank * 3:(rank + 1) * 3]).requires_grad_(True)
  (ga, gp), scale = losses.gather_global_batch(a, p)
  loss = (oracle.npairs_loss(ga, gp) + oracle.npairs_loss(gp, ga)) * scale
  loss.backward()
  np.save(os.path.join(out_dir, 'gather%d.npy' % rank), {'loss': float(loss.detach()), 'scale': scale, 'shape': tuple(ga.shape),
                                                         'da': a.grad.numpy(), 'dp': p.grad.numpy()}, allow_pickle=True)
  dist.barrier()
  dist.destroy_process_group()


def test_global_negatives_gather_two_ranks(tmp_path):
  """SURVEY 8(e) opt-in for Grasp2Vec: the n-pairs loss over the all-gathered embeddings.  Every rank computes the same
  global loss; after the step's 1 / world gradient scaling the per-sample gradients equal those of ONE process holding
  the whole batch."""
  from oracle import grasp2vec as oracle
  world, port = 2, _free_port()
  mp.spawn(_gather_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
  res = [np.load(os.path.join(str(tmp_path), 'gather%d.npy' % r), allow_pickle=True).item() for r in range(world)]
  rng = np.random.RandomState(0)
  a = torch.from_numpy(rng.standard_normal((world * 3, 8)) * 0.5).requires_grad_(True)
  p = torch.from_numpy(rng.standard_normal((world * 3, 8)) * 0.5).requires_grad_(True)
  single = oracle.npairs_loss(a, p) + oracle.npairs_loss(p, a)
  single.backward()
  for r, out in enumerate(res):
    assert out['shape'] == (world * 3, 8) and out['scale'] == float(world)
    np.testing.assert_allclose(out['loss'] / world, float(single.detach()), rtol=1e-10)
    np.testing.assert_allclose(out['da'] / world, a.grad.numpy()[r * 3:(r + 1) * 3], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(out['dp'] / world, p.grad.numpy()[r * 3:(r + 1) * 3], rtol=1e-9, atol=1e-12)
