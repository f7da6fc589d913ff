"""The warp-specialised fixed-step flow kernel
(principia_b200/csrc/pb_kernels_massless_specialized.cuh; opt-in,
PB_SRKN_SPECIALIZED=1): the same parity tests as the default kernel, in a
process of their own because the library reads the switch once."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def test_specialised_kernel_is_bit_identical():
  """Full-size LEO batch (every block completes in the specialised kernel) and
  the mixed batch (the blocks with a lunar probe are given up and flowed by the
  general kernel): subsets bit-equal to the oracle, samples included."""
  environment = dict(os.environ, PB_SRKN_SPECIALIZED='1')
  completed = subprocess.run(
      [sys.executable, '-m', 'pytest', '-q', '-x', '-m', 'gpu',
       os.path.join(HERE, 'test_gpu_massless.py'), '-k',
       'full_size_flow_properties or mixed_batch or collision_status'],
      env=environment, capture_output=True, text=True, timeout=900)
  assert completed.returncode == 0, completed.stdout[-3000:] + completed.stderr[-2000:]
  assert ' passed' in completed.stdout
