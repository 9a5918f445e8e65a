import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'gpu_next: GPU test of a code path written after the round\'s GPU budget was '
                            'spent -- not yet run on hardware, opt-in product path; run with -m gpu_next, promote to '
                            'gpu once green (tools/gpu_round2_first.sh)')


    config.addinivalue_line('markers', 'product_precision: run with the library\'s default precision policy (UNet inference '
                            'in the fp8-corrected operand format) instead of the three-pass policy the tight parity '
                            'thresholds of the other tests are written for')


@pytest.fixture(autouse=True)
def _precision_policy(request):
    """Parity thresholds of the suite (TIGHT = 5e-5 ...) are those of the fp32-faithful three-pass products; tests of the
    fp8-corrected inference format select it explicitly (tests/test_fp8c_gpu.py) or carry the product_precision marker."""
    from slotdiffusion_b200 import ops
    saved = (ops._PASSES, ops._UNET_INFERENCE)
    if request.node.get_closest_marker('product_precision') is None:
        ops.set_precision('fp32')
    yield
    ops._PASSES, ops._UNET_INFERENCE = saved


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for it in items:
        if 'gpu' in it.keywords or 'gpu_next' in it.keywords:
            it.add_marker(skip)
