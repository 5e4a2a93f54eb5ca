import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


if os.environ.get('XRVIO_EMU') == '1':
    # TEST INFRASTRUCTURE: run the "gpu" tests against the product's kernels compiled for the CPU by the SIMT emulator in
    # tests/cuemu (debugging aid for machines without a GPU; tests/test_emulated.py drives it).  The product never does this.
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'cuemu'))
    import build_emu
    from xrslam_b200 import capi as _capi
    _capi.LIB_PATH = build_emu.build()
    os.environ.setdefault('XRVIO_LANES', '1')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run with -m gpu on the B200 box)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def xrvio():
    """The C-ABI library; fails loudly if it is not built or no GPU is present (no CPU fallback)."""
    from xrslam_b200 import capi
    n = capi.lib.xrvio_device_count()
    assert n > 0, 'no CUDA device: ' + capi.lib.xrvio_last_error().decode()
    return capi
