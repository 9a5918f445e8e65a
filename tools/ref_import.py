"""Import the UNMODIFIED reference (Wuziyi616/SlotDiffusion) with stubs for its un-vendored dependencies.

Users: tools/make_golden.py (build container, /root/reference), and -- through the copy under baseline/_ref that
travels to the GPU box -- tests/test_reference_dropin_gpu.py, tools/ref_gpu_bar.py and bench.py's reference /
gpu_baseline legs.  The product never imports this file.  Recipe follows SURVEY.md Appendix B.
"""
import importlib
import os
import sys
import types
import warnings

STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_stubs')
_BOX_COPY = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'baseline', '_ref')


def _find_root():
    """/root/reference in the build container; on the GPU box only baseline/_ref exists (the unmodified python tree put
    there by baseline/install_ref.sh: git-ignored, travels with the gpurun snapshot)."""
    for c in (os.environ.get('SDB_REFERENCE_ROOT'), '/root/reference', _BOX_COPY):
        if c and os.path.isdir(os.path.join(c, 'slotdiffusion')):
            return c
    return '/root/reference'


REF_ROOT = _find_root()


def box_copy_available():
    return os.path.isdir(os.path.join(_BOX_COPY, 'slotdiffusion'))


def use_box_copy():
    """Point this module at baseline/_ref (what the -m gpu tests and bench.py use: nothing there may read /root/reference)."""
    global REF_ROOT
    REF_ROOT = _BOX_COPY


def setup():
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f'reference not found at {REF_ROOT}')
    warnings.filterwarnings('ignore')
    if STUBS not in sys.path:
        sys.path.insert(0, STUBS)
    if REF_ROOT not in sys.path:
        sys.path.insert(1, REF_ROOT)
    # transformers 5.x no longer exports ViTFeatureExtractor (dino.py:6)
    fake = types.ModuleType('transformers')
    fake.ViTFeatureExtractor = object
    fake.ViTModel = object
    sys.modules['transformers'] = fake
    if 'wandb' not in sys.modules:
        try:
            import wandb  # noqa
        except Exception:
            sys.modules['wandb'] = types.ModuleType('wandb')


def load_params(task, cfg_relpath):
    """Like scripts/train.py:103-108."""
    setup()
    cfg = os.path.join(REF_ROOT, 'slotdiffusion', task, 'configs', cfg_relpath)
    d, f = os.path.split(cfg)
    sys.path.insert(0, d)
    mod = importlib.import_module(f[:-3])
    sys.path.pop(0)
    return mod.SlotAttentionParams()


def fresh_params(task, cfg_relpath):
    """A params object from a FRESH class: build_model() pops entries from the class-level dicts, so every build needs
    its own copy of the config module."""
    import runpy
    setup()
    return runpy.run_path(os.path.join(REF_ROOT, 'slotdiffusion', task, 'configs', cfg_relpath))['SlotAttentionParams']()


def img_models():
    setup()
    return importlib.import_module('slotdiffusion.img_based.models')


def video_models():
    setup()
    return importlib.import_module('slotdiffusion.video_based.models')
