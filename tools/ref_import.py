"""Import the reference (Wuziyi616/SlotDiffusion) in the build container.

Only tools/make_golden.py and ad-hoc probes use this.  /root/reference does not
exist on the GPU box, so nothing under tests/, bench.py or the product imports
this file.  Recipe follows SURVEY.md Appendix B.
"""
import importlib
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get('SDB_REFERENCE_ROOT', '/root/reference')
STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_stubs')


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


def img_models():
    setup()
    return importlib.import_module('slotdiffusion.img_based.models')


def video_models():
    setup()
    return importlib.import_module('slotdiffusion.video_based.models')
