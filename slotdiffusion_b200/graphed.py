"""CUDA-graph capture of the TRAINING forward / backward of the B200 modules, for eager training loops.

The reference trains with an eager Python loop (nerv: forward -> loss.backward() -> optimizer.step(), scripts/train.py);
the kernel schedules of the modules here cost ~13 ms (forward) + ~23 ms (backward) of Python per UNet step when they are
issued launch by launch -- about as much as the device time.  `enable(module)` makes the module's
gradient-enabled forward run through torch.cuda.make_graphed_callables: per input-shape signature, the forward schedule
and the backward schedule are each captured ONCE into a CUDA graph (after the usual warm-up iterations) and replayed from
then on, still as a node of the caller's autograd graph -- the caller's loop, loss and optimizer stay untouched.

What makes the schedules capturable: every entry point of libsdb200 only enqueues work on the current stream, no host
read-back, buffers come from torch's graph-private pool.  Two things that are normally decided on the host are moved into
the captured work: parameters are re-packed inside the captured forward (ops.WeightCache.nocache: a replay after
optimizer.step() sees the new values), and the dropout step counter lives in device memory and is advanced by the captured
forward (ops.dropout_step_counter) so every replay draws new masks.

dropin.install(graph=True) enables this for every B200 module the reference constructs.
"""
import torch
from torch import nn


class _Core(nn.Module):
    """What make_graphed_callables captures: owner's raw training forward over positional tensors; reports the owner's
    parameters as its own so that their gradients are outputs of the captured backward."""

    def __init__(self, owner, fn, caches):
        super().__init__()
        object.__setattr__(self, '_owner', owner)
        object.__setattr__(self, '_fn', fn)
        object.__setattr__(self, '_caches', caches)

    def parameters(self, recurse=True):
        return self._owner.parameters(recurse)

    def forward(self, *tensors):
        from . import ops
        dev = tensors[0].device
        ops.dropout_step_counter(dev).add_(1)          # inside the capture: a new dropout mask stream per replay
        for wc in self._caches():
            # (re)pack the weights inside the training schedules -- forward AND backward (transposed / rotated operands of
            # the dgrad GEMMs) -- so that the captured graphs contain the packing kernels.  Honoured only inside
            # ops.training_scope(): the no-grad inference paths of the same module keep their cache.
            wc.nocache = True
        return self._fn(*tensors)


class GraphedTraining:
    """Per-shape cache of graphed training callables of one module."""

    def __init__(self, owner, fn, caches):
        self.owner, self.fn, self.caches = owner, fn, caches
        self.graphs = {}

    def __call__(self, *tensors):
        key = (self.owner.training,) + tuple((tuple(t.shape), t.dtype, bool(t.requires_grad)) for t in tensors)
        g = self.graphs.get(key)
        if g is None:
            core = _Core(self.owner, self.fn, self.caches)
            core.train(self.owner.training)
            sample = tuple(t.detach().clone().requires_grad_(t.requires_grad) for t in tensors)
            g = torch.cuda.make_graphed_callables(core, sample, allow_unused_input=True)
            self.graphs[key] = g
        return g(*tensors)

    def clear(self):
        self.graphs.clear()


def enable(module):
    """Route the gradient-enabled forward of a slotdiffusion_b200 module through captured CUDA graphs (idempotent).
    Supported: slot_attention.SlotAttention / SlotAttentionWMask, unet.UNetModel, resnet.ResNet."""
    module.__dict__['_sdb_graphed'] = True
    return module


def disable(module):
    module.__dict__.pop('_sdb_graphed', None)
    g = module.__dict__.pop('_sdb_graphs', None)
    if g is not None:
        g.clear()
    return module


def graphs_of(module, fn, caches):
    """The module's GraphedTraining cache (created on first use; dropped by deepcopy / pickle with the other derived state)."""
    g = module.__dict__.get('_sdb_graphs')
    if g is None:
        g = module.__dict__['_sdb_graphs'] = GraphedTraining(module, fn, caches)
    return g


def enabled(module):
    return bool(module.__dict__.get('_sdb_graphed', False)) and not torch.cuda.is_current_stream_capturing()
