"""bench.py contract checks that run without a GPU: the reference arm (the reference's CPU path: the unmodified reference
from baseline/_ref when that copy is present -- kind "reference" -- else the oracle port) prints
ONE JSON line with the agreed keys, and the product arm refuses to run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, cwd=ROOT, timeout=580)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference'
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['metric'] == 'dpm_solver_denoise_steps_per_sec' and d['unit'] == 'sample-steps/s'
    have_ref = os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'slotdiffusion'))
    assert d['value'] > 0 and d['cpu_baseline']['kind'] == ('reference' if have_ref else 'port')
    assert d['cpu_baseline']['cores'] >= 1 and d['config']['per_gpu_batch'] == 4
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['vs_baseline'] is None


def test_product_modules_refuse_cpu_tensors():
    """no CPU / eager fallback on the product path"""
    import torch
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    mod = SlotAttentionWMask(64, 1, 3, 64, 128)
    with pytest.raises(RuntimeError):
        mod(torch.randn(1, 16, 64), torch.randn(1, 3, 64))
