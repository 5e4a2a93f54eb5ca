"""The product's CUDA kernels executed on the CPU by the test-only SIMT emulator (tests/cuemu): the operator-level
parity tests of tests/test_gpu_ops.py, the undistortion tests and the large-window solver tests run against libxrvio_emu.so -- the unmodified .cu
sources compiled by g++ with fibers standing in for CUDA threads.  This is a debugging aid for machines without a GPU
(a kernel change can be checked against the oracle before it is measured on a B200); it is not a product path: the
shipped library is built by nvcc only and refuses to run without a CUDA device (test_no_cpu_fallback_without_gpu).
The session-level tests also run under the emulator (XRVIO_EMU=1 python -m pytest tests -m gpu, ~10 min); only the
fast operator-level ones are part of the CPU suite."""
import os
import subprocess
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(900)
def test_operator_parity_on_the_emulator():
    env = dict(os.environ, XRVIO_EMU='1')
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_gpu_ops.py'), os.path.join(ROOT, 'tests', 'test_gpu_zundistort.py'),
                        os.path.join(ROOT, 'tests', 'test_gpu_zz_large_window.py'),
                        '-q', '-m', 'gpu', '-x', '-p', 'no:cacheprovider'], capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert ' passed' in r.stdout and 'failed' not in r.stdout
