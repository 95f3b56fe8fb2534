#!/usr/bin/env python
"""bench.py - QT-Opt replay transitions/sec on B200 (BASELINE.json metric), one JSON line.

  python bench.py --gpus 1 --steps K --warmup W             # this engine
  python bench.py --impl reference --steps K --warmup W     # the reference path's CPU restatement
  torchrun ... bench.py --gpus N ...                        # data parallel, one rank per GPU

A step = one pass of the hot path over one replay batch: crop/convert of the uint8 512x640 frames
-> ResNet-50 Q-critic forward -> log loss -> backward -> (NCCL all-reduce) -> fused optimizer+EMA.
`value` times that with the batch resident in HBM; `e2e` times it through the public step call with
the batch in pinned host memory (H2D inside the timed region) and the loss read back every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = 'qtopt_replay_transitions_per_sec'
# Algorithmic work per transition (SURVEY.md 8d / BASELINE.md 2; 2*MACs, train = 3x forward).
TRAIN_GFLOP_PER_TRANSITION = {'resnet50': 110.9, 'grasping44': 26.4}


def parse_args():
  p = argparse.ArgumentParser()
  p.add_argument('--gpus', type=int, default=1)
  p.add_argument('--steps', type=int, default=8)
  p.add_argument('--warmup', type=int, default=3)
  p.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  p.add_argument('--model', default='resnet50', choices=['resnet50', 'grasping44'])
  p.add_argument('--batch', type=int, default=512, help='transitions per GPU per step')
  p.add_argument('--cpu-batch', type=int, default=16)
  p.add_argument('--no-cpu-baseline', action='store_true')
  p.add_argument('--no-e2e', action='store_true')
  p.add_argument('--no-cem', action='store_true')
  p.add_argument('--cem-batch', type=int, default=64, help='transitions per GPU for the CEM measurement')
  return p.parse_args()


# ---------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md recipe), sampled DURING the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
  QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
           'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
           'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index, self.proc, self.lines = index, None, []

  def start(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                    '--format=csv,noheader,nounits', '-lms', '200'],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except subprocess.TimeoutExpired:
      self.proc.kill()
    sm, smax, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for line in self.lines:
      f = [x.strip() for x in line.split(',')]
      if len(f) < 9:
        continue
      try:
        sm.append(float(f[1]))
        smax.append(float(f[2]))
      except ValueError:
        continue
      for name, v in zip(names, f[5:9]):
        if v.lower().startswith('active'):
          reasons.add(name)
    return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
            'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      d = json.load(f)
    return {'tflops': d.get('bf16_tflops_sustained', 1400.0), 'tflops_burst': d.get('bf16_tflops', 1590.0),
            'hbm_gbs': d.get('hbm_gbs', 6650.0), 'source': 'measured'}
  return {'tflops': 1400.0, 'tflops_burst': 1590.0, 'hbm_gbs': 6650.0, 'source': 'fallback'}


# ---------------------------------------------------------------------------------------------
# CPU baseline: the oracle restatement of the reference step (bench `cpu_baseline` / --impl reference)
# ---------------------------------------------------------------------------------------------
def usable_host_threads():
  """Host threads this process can really run: the affinity mask capped by the cgroup CPU quota (a
  128-core box with a 16-CPU quota runs a 128-thread torch step 70x slower than a 16-thread one)."""
  n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
  try:
    quota, period = open('/sys/fs/cgroup/cpu.max').read().split()
    if quota != 'max':
      n = min(n, max(1, int(int(quota) / int(period))))
  except (OSError, ValueError):
    try:
      quota = int(open('/sys/fs/cgroup/cpu/cpu.cfs_quota_us').read())
      period = int(open('/sys/fs/cgroup/cpu/cpu.cfs_period_us').read())
      if quota > 0:
        n = min(n, max(1, quota // period))
    except (OSError, ValueError):
      pass
  return n


def cpu_reference_step_rate(model, batch, steps, warmup, size=472):
  """Transitions/s of the torch-CPU fp32 restatement of the reference train step (forward, log loss
  + l2, backward, momentum update) on all host cores.  Executes oracle/ - allowed only here."""
  from oracle import qtopt_networks, resnet as oracle_resnet, tf_ops
  threads = usable_host_threads()
  torch.set_num_threads(threads)
  rng = np.random.RandomState(0)
  img = torch.from_numpy(rng.uniform(0, 1, (batch, size, size, 3)).astype(np.float32))
  grasp = torch.from_numpy(rng.uniform(-1, 1, (batch, 10)).astype(np.float32))
  reward = torch.from_numpy((rng.uniform(size=(batch, 1)) < 0.3).astype(np.float32))
  if model == 'grasping44':
    variables = qtopt_networks.to_torch(qtopt_networks.init_variables(0))
    fwd = lambda: qtopt_networks.model(variables, img, grasp, True)
  else:
    variables = {}
    with torch.no_grad():
      oracle_resnet.critic(variables, img[:1], grasp[:1], False, rng=np.random.RandomState(0))
    for k, v in variables.items():
      v.requires_grad_(not (k.endswith('moving_mean') or k.endswith('moving_variance')))
    fwd = lambda: oracle_resnet.critic(variables, img, grasp, True)
  params = [v for v in variables.values() if v.requires_grad]
  momentum = [torch.zeros_like(p) for p in params]

  def step():
    logits = fwd()
    loss = tf_ops.log_loss(reward, torch.sigmoid(logits))
    loss = loss + sum(tf_ops.l2_regularizer(7e-5, v) for k, v in variables.items()
                      if k.endswith('/weights') or k.endswith('/kernel'))
    grads = torch.autograd.grad(loss, params)
    with torch.no_grad():
      for p, g, m in zip(params, grads, momentum):
        m.mul_(0.9).add_(g)
        p.sub_(1e-4 * m)
    return float(loss.detach())

  for _ in range(warmup):
    step()
  t0 = time.perf_counter()
  for _ in range(steps):
    step()
  dt = time.perf_counter() - t0
  return batch * steps / dt, threads, dt / steps


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  steps, warmup = max(1, min(args.steps, 6)), max(0, min(args.warmup, 1))
  rate, threads, sec = cpu_reference_step_rate(args.model, args.cpu_batch, steps, warmup)
  sample = '%d timed steps (+%d warm-up) of batch %d, torch-CPU fp32 restatement of the reference step' % (
      steps, warmup, args.cpu_batch)
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'transitions/s', 'n_gpus': args.gpus,
      'steps': steps, 'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      # the same workload as the engine arm; the bounded CPU sample is described in cpu_baseline.sample
      'config': workload_config(args, per_gpu_batch=args.batch),
      'cpu_baseline': {'value': rate, 'unit': 'transitions/s', 'cores': threads, 'kind': 'port', 'sample': sample},
      'e2e': {'value': rate, 'unit': 'transitions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }
  print(json.dumps(line))


def workload_config(args, per_gpu_batch):
  name = ('QT-Opt ResNet-50 Q-critic train step' if args.model == 'resnet50' else
          'QT-Opt Grasping44 Q-critic train step')
  return {'workload': name + ', synthetic 512x640 uint8 replay frames -> 472x472 crop',
          'critic': args.model, 'per_gpu_batch': per_gpu_batch, 'global_batch': per_gpu_batch * args.gpus,
          'image': '472x472x3', 'optimizer': 'momentum+EMA', 'parallelism': 'dp%d' % args.gpus,
          'l2_flush': 'inputs (503 MB/step at batch 512) and activations exceed the 126 MB L2'}


# ---------------------------------------------------------------------------------------------
# the engine
# ---------------------------------------------------------------------------------------------
def run_b200(args):
  import torch.distributed as dist
  from tensor2robot_b200 import _lib, engine, nn
  from tensor2robot_b200.models import optimizers
  from tensor2robot_b200.research.qtopt import networks, resnet_critic

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if world != args.gpus:
    if world == 1 and args.gpus > 1:
      raise SystemExit('launch with torchrun --nproc-per-node %d for --gpus %d' % (args.gpus, args.gpus))
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)

  b = args.batch
  critic = resnet_critic.ResNet50QCritic() if args.model == 'resnet50' else \
      networks.Grasping44E2EOpenCloseTerminateGripperStatusHeightToBottom()
  # research/qtopt/optimizer_builder.py defaults: momentum 0.9, staircase decay, EMA 0.9999
  lr = optimizers.create_exp_decaying_learning_rate(1e-4, int(3e6 / 32 * 2), 0.999, True)
  opt = optimizers.MovingAverageOptimizer(optimizers.MomentumOptimizer(lr, 0.9), 0.9999)
  step = engine.CriticTrainStep(critic, opt, device=dev, seed=0, world_size=world, rank=rank)

  g = torch.Generator(device=dev)
  g.manual_seed(1234 + rank)
  n_sets = 2
  dev_batches = []
  for i in range(n_sets):
    dev_batches.append((torch.randint(0, 256, (b, 512, 640, 3), dtype=torch.uint8, device=dev, generator=g),
                        torch.rand((b, 10), device=dev, generator=g) * 2 - 1,
                        (torch.rand((b, 1), device=dev, generator=g) < 0.3).float()))
  step.build(*dev_batches[0][:2])

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # ---- device-resident timing ----
  for i in range(args.warmup):
    step.step(*dev_batches[i % n_sets])
  barrier()
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  nn.PROFILE = []
  launches0 = _lib.launch_count()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ev0.record()
  for i in range(args.steps):
    loss = step.step(*dev_batches[i % n_sets])
  ev1.record()
  barrier()
  launches = _lib.launch_count() - launches0
  prof, nn.PROFILE = nn.PROFILE, None
  clocks = sampler.stop() if rank == 0 else None
  elapsed_ms = ev0.elapsed_time(ev1)
  if world > 1:
    t = torch.tensor([elapsed_ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
  loss_value = float(loss)

  # ---- roofline of the dominant kernel family (tcgen05 convolutions), live CUDA-event timings ----
  by_kind, by_shape = {}, {}
  for tag, flops, e0, e1 in prof:
    ms = e0.elapsed_time(e1)
    for table, key in ((by_kind, tag.split('|')[0]), (by_shape, tag)):
      d = table.setdefault(key, [0.0, 0.0, 0])
      d[0] += flops
      d[1] += ms
      d[2] += 1
  if os.environ.get('T2R_BENCH_DETAIL') and rank == 0:
    for key, v in sorted(by_shape.items(), key=lambda kv: -kv[1][1]):
      sys.stderr.write('%-48s n=%3d  %8.3f ms/step  %7.1f TFLOP/s\n' % (
          key, v[2] // args.steps, v[1] / args.steps, v[0] / max(v[1], 1e-9) / 1e9))
  peaks = measured_peaks()
  conv_ms = sum(v[1] for v in by_kind.values())
  conv_flops = sum(v[0] for v in by_kind.values())
  dominant = max(by_kind.items(), key=lambda kv: kv[1][1])[0] if by_kind else None
  kinds = {k: {'launches': v[2], 'ms': v[1], 'tflops': v[0] / max(v[1], 1e-9) / 1e9} for k, v in by_kind.items()}
  achieved = conv_flops / max(conv_ms, 1e-9) / 1e9
  roofline = {
      'bound': 'tensor',
      'kernel': 'conv_igemm_kernel / conv_igemm_tma_kernel / conv_halo_kernel / conv_wgrad_kernel (tcgen05 implicit GEMM)',
      'achieved': achieved, 'peak': peaks['tflops'], 'unit': 'TFLOP/s', 'frac': achieved / peaks['tflops'],
      'peak_source': peaks['source'] + ' sustained cuBLAS bf16 (kernel timed inside a long step)',
      # ncu --set full, one launch of conv_igemm_kernel<256> (3x3 256->256 fprop at B=512): dram read + write
      # bytes = 435 MB against 472 MB algorithmic (profiles/r01_ncu_summary.md, section 5)
      'traffic': 434.9e6 if args.model == 'resnet50' else None,
      'traffic_kernel': 'conv_igemm_kernel<256> 512x30x30x256->256 k3 (ncu, profiles/r01g_igemm256_3x3_256.ncu-rep)',
      'dominant': dominant, 'by_kind': kinds,
      'share_of_step': conv_ms / max(elapsed_ms, 1e-9),
  }

  # ---- end to end through the public step call: pinned host batch -> H2D -> step -> loss D2H ----
  e2e = None
  if not args.no_e2e:
    host = [tuple(t.cpu().pin_memory() for t in bt) for bt in dev_batches]
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    slots = [tuple(torch.empty_like(t, device=dev) for t in host[0]) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)

    def stage(i):   # H2D of batch i on the copy stream, overlapping the previous step's compute
      with torch.cuda.stream(copy_stream):
        for dst, src in zip(slots[i % 2], host[i % n_sets]):
          dst.copy_(src, non_blocking=True)
        done = torch.cuda.Event()
        done.record(copy_stream)
      return done

    def e2e_loop(n):
      losses = []
      ready = stage(0)
      for i in range(n):
        torch.cuda.current_stream().wait_event(ready)
        if i + 1 < n:
          ready = stage(i + 1)
        l = step.step(*slots[i % 2])
        losses.append(l.to('cpu', non_blocking=False))   # device -> host read of the step result
      return losses

    e2e_loop(max(2, args.warmup))
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
      t = torch.tensor([e2e_s], device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      e2e_s = float(t.item())
    e2e = {'value': b * world * args.steps / e2e_s, 'unit': 'transitions/s', 'h2d_bytes_per_step': h2d,
           'd2h_bytes_per_step': 4, 'ms_per_step': e2e_s / args.steps * 1e3}

  # ---- CEM action maximisation (BASELINE metric "CEM Q-evals/sec"): 64 samples x 2 iterations per
  # transition against the staged state features, + Bellman target ----
  cem_line = None
  if not args.no_cem:
    cb = min(b, args.cem_batch)
    cem = engine.CEMTargetComputer(critic, step.vs, action_size=10, cem_samples=64, cem_iters=2, num_elites=10,
                                   seed=rank, chunk=16 if args.model == 'resnet50' else None)
    frames = dev_batches[0][0][:cb]
    reward, done = dev_batches[0][2][:cb, 0].contiguous(), torch.zeros(cb, device=dev)

    def cem_once():
      x = step.preprocess(frames, training=False)
      _, max_q, _ = cem.maximize(x)
      return cem.bellman_target(reward, done, max_q, 0.9)

    cem_once()
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(2):
      target = cem_once()
    c1.record()
    barrier()
    cem_ms = c0.elapsed_time(c1) / 2
    if world > 1:
      t = torch.tensor([cem_ms], device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      cem_ms = float(t.item())
    cem_line = {'q_evals_per_sec': cb * world * 64 * 2 * 1000.0 / cem_ms, 'transitions_per_sec': cb * world * 1000.0 / cem_ms,
                'per_gpu_batch': cb, 'samples': 64, 'iterations': 2, 'elites': 10, 'ms': cem_ms,
                'target_mean': float(target.mean())}

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return
  cpu = None
  if world == 1 and not args.no_cpu_baseline:
    rate, threads, sec = cpu_reference_step_rate(args.model, args.cpu_batch, 3, 1)
    cpu = {'value': rate, 'unit': 'transitions/s', 'cores': threads, 'kind': 'port',
           'sample': '3 timed steps (+1 warm-up) of batch %d, torch-CPU fp32 restatement of the reference step '
                     '(%.1f s/step)' % (args.cpu_batch, sec)}
  ms_per_step = elapsed_ms / args.steps
  value = b * world * 1000.0 / ms_per_step
  line = {
      'metric': METRIC, 'value': value, 'unit': 'transitions/s', 'n_gpus': world, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
      'config': workload_config(args, b), 'clocks': clocks, 'gpu_launches': int(launches),
      'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'cem': cem_line, 'loss': loss_value,
      'model_tflops': value * TRAIN_GFLOP_PER_TRANSITION[args.model] / 1e3,
      'peak_mem_gb': torch.cuda.max_memory_allocated(dev) / 1e9,
  }
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def main():
  args = parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_b200(args)


if __name__ == '__main__':
  main()
