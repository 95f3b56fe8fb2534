import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
  import torch
  if torch.cuda.is_available():
    return
  skip = pytest.mark.skip(reason='no CUDA device in this container')
  for item in items:
    if 'gpu' in item.keywords:
      item.add_marker(skip)


_GPU_NODES = set()


def pytest_itemcollected(item):
  if 'gpu' in item.keywords:
    _GPU_NODES.add(item.nodeid)


def pytest_runtest_logreport(report):
  """Every GPU parity test prints the errors it measured; keep them: one JSON line per test (node id, outcome, the
  printed measurements) appended to gpurun_out/parity_measurements.jsonl, which travels back from the GPU box and is
  copied to profiles/ for the record."""
  if report.when != 'call' or report.nodeid not in _GPU_NODES or report.outcome == 'skipped':
    return
  import json
  lines = [l for l in (report.capstdout or '').splitlines() if l.strip()]
  path = os.path.join(ROOT, 'gpurun_out', 'parity_measurements.jsonl')
  try:
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, 'a') as f:
      record = {'test': report.nodeid, 'outcome': report.outcome, 'seconds': round(report.duration, 3), 'measured': lines}
      if report.outcome == 'failed':
        record['error'] = str(report.longrepr)[-1500:]
      f.write(json.dumps(record) + '\n')
  except OSError:
    pass
