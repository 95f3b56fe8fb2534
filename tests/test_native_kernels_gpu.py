"""Runs the stand-alone native harness (tests/native/test_kernels.cu: every kernel against naive CPU loops, no
Python / torch in the process) as part of the GPU suite, once in the default configuration and once with the
opt-in fused batch-norm-backward epilogue, and keeps the log under gpurun_out/."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINARY = os.path.join(ROOT, 'tests', 'native', 'test_kernels')


def _run(args, env_extra, log_name):
  if not os.path.exists(BINARY):
    pytest.fail('tests/native/test_kernels is not built: run __graft_entry__.build()')
  env = dict(os.environ, **env_extra)
  out = subprocess.run([BINARY] + args, cwd=os.path.dirname(BINARY), env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
  try:
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', log_name), 'w') as f:
      f.write(out.stdout)
  except OSError:
    pass
  failures = [l for l in out.stdout.splitlines() if l.startswith(('[FAIL]', '[FATAL]'))]
  assert out.returncode == 0 and not failures, '\n'.join(failures[:20] + out.stdout.splitlines()[-3:])
  return out.stdout


def test_native_harness_all_kernels():
  text = _run([], {}, 'r02_native_kernel_tests.log')
  assert 'SUMMARY' in text and 'fail=0' in text


def test_native_harness_fused_bn_backward_epilogue():
  text = _run(['bnfuse'], {'T2R_BNBWD_EPI': '1'}, 'r02_native_bnfuse_epilogue.log')
  assert 'dgrad g (masked)' in text and 'fail=0' in text
