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


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for it in items:
        if 'gpu' in it.keywords or 'gpu_next' in it.keywords:
            it.add_marker(skip)
