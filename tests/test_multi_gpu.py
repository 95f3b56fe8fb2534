"""Two ranks over NCCL on two GPUs of one box (skipped on a single-GPU box; host-side logic of the same path is covered
on CPU by tests/test_data_parallel_cpu.py).  SURVEY 8(e): one process per GPU, rank 0 broadcasts the initial
parameters, the flat gradient buffer is all-reduced in buckets while the backward pass runs, 1 / world is folded into
the optimizer kernel.

Check: both ranks feed the SAME batch, so the averaged gradient equals each replica's own gradient and the replicas
must (a) stay bit-identical to each other after every step and (b) match a single-process run of the same steps up to
the summation order of the weight-gradient atomics."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _run(rank, world, port, out_dir):
  import torch.distributed as dist
  from tensor2robot_b200 import engine
  from tensor2robot_b200.models import optimizers
  from tensor2robot_b200.research.qtopt import networks
  torch.cuda.set_device(rank)
  if world > 1:
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
  rng = np.random.RandomState(0)
  frames = torch.from_numpy(rng.randint(0, 256, (8, 512, 640, 3)).astype(np.uint8)).cuda()
  actions = torch.from_numpy(rng.uniform(-1, 1, (8, 10)).astype(np.float32)).cuda()
  reward = torch.from_numpy((rng.uniform(size=(8, 1)) < 0.4).astype(np.float32)).cuda()
  critic = networks.Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom()
  opt = optimizers.MovingAverageOptimizer(optimizers.MomentumOptimizer(1e-4, 0.9), 0.99)
  # seed differs per rank ON PURPOSE: build() must replace every replica's initial values by rank 0's
  step = engine.CriticTrainStep(critic, opt, device='cuda:%d' % rank, seed=rank, world_size=world, rank=0)
  step._rng = np.random.RandomState(123)      # the same random crops on every replica and in the single-process run
  step.build(frames, actions)
  losses = []
  for _ in range(3):
    losses.append(float(step.step(frames, actions, reward)))
  early = step.reducer.launched_early if step.reducer is not None else 0
  flat = step.vs.flat.detach().clone()
  same = True
  if world > 1:
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
  np.save(os.path.join(out_dir, 'w%d_rank%d.npy' % (world, rank)),
          {'flat': flat.cpu().numpy(), 'losses': losses, 'replicas_identical': same, 'buckets_early': early,
           'n_buckets': len(step.reducer.buckets) if step.reducer is not None else 0}, allow_pickle=True)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_gpu_replicas_match_single_process(tmp_path):
  import torch.multiprocessing as mp
  mp.spawn(_run, args=(1, 0, str(tmp_path)), nprocs=1, join=True)
  mp.spawn(_run, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
  single = np.load(str(tmp_path / 'w1_rank0.npy'), allow_pickle=True).item()
  ranks = [np.load(str(tmp_path / ('w2_rank%d.npy' % r)), allow_pickle=True).item() for r in range(2)]
  assert all(r['replicas_identical'] for r in ranks)
  assert ranks[0]['n_buckets'] >= 2 and ranks[0]['buckets_early'] >= 1      # the all-reduce started before backward ended
  for r in ranks:
    # a training-mode BN network amplifies the summation-order noise of the weight-gradient atomics from step to step
    np.testing.assert_allclose(r['losses'], single['losses'], rtol=5e-3)
    err = np.linalg.norm(r['flat'] - single['flat']) / np.linalg.norm(single['flat'])
    assert err < 1e-3, err
