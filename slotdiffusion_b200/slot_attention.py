"""Drop-in Slot Attention modules backed by libsdb200 (hand-written sm_100a kernels).

Mirrors the reference module interface exactly -- constructor arguments, attribute / parameter names and
shapes (so reference checkpoints load unchanged) and forward() signatures:
  * SlotAttention        /root/reference/slotdiffusion/img_based/models/slot_attention.py:15-112
                         (duplicate: video_based/models/savi.py:17-114)
  * SlotAttentionWMask   img_based/models/sa_diffusion.py:9-70 (video_based/models/savi_diffusion.py:10-71)
The nn.Module objects only own the parameters; all arithmetic runs in the CUDA library.  There is no
PyTorch fallback: calling forward() on CPU tensors raises.
"""
import torch
from torch import nn

from . import ops
from .autograd import slot_attention_apply


class SlotAttention(nn.Module):
    """Slot attention module that iteratively performs cross-attention (same ctor as the reference)."""

    def __init__(self, in_features, num_iterations, num_slots, slot_size, mlp_hidden_size, eps=1e-6):
        super().__init__()
        self.in_features = in_features
        self.num_iterations = num_iterations
        self.num_slots = num_slots
        self.slot_size = slot_size
        self.mlp_hidden_size = mlp_hidden_size
        self.eps = eps
        self.attn_scale = self.slot_size ** -0.5

        self.norm_inputs = nn.LayerNorm(self.in_features)
        self.project_q = nn.Sequential(
            nn.LayerNorm(self.slot_size),
            nn.Linear(self.slot_size, self.slot_size, bias=False),
        )
        self.project_k = nn.Linear(in_features, self.slot_size, bias=False)
        self.project_v = nn.Linear(in_features, self.slot_size, bias=False)
        self.gru = nn.GRUCell(self.slot_size, self.slot_size)
        self.mlp = nn.Sequential(
            nn.LayerNorm(self.slot_size),
            nn.Linear(self.slot_size, self.mlp_hidden_size),
            nn.ReLU(),
            nn.Linear(self.mlp_hidden_size, self.slot_size),
        )
        self._wcache = ops.WeightCache()

    def __getstate__(self):          # deepcopy / pickle: the packed-weight cache is derived state
        d = self.__dict__.copy()
        d.pop('_wcache', None)
        d.pop('_gradbuf', None)      # gradient layout is keyed by id(parameter) of THIS instance
        d.pop('_sdb_graphs', None)   # captured training graphs address this instance's parameters
        return d

    def __setstate__(self, d):
        super().__setstate__(d)
        self._wcache = ops.WeightCache()

    def invalidate_caches(self):
        """Drop the packed / folded weights (needed after parameter updates made through `p.data`, which change neither
        data_ptr nor _version -- the two things the cache watches)."""
        self._wcache._c.clear()

    def _run(self, inputs, slots, want_mask):
        if not inputs.is_cuda:
            raise RuntimeError('slotdiffusion_b200.SlotAttention runs on CUDA (sm_100a) only; no CPU fallback')
        assert slots.dim() == 3 and inputs.dim() == 3
        from . import graphed
        if graphed.enabled(self) and torch.is_grad_enabled() and (
                inputs.requires_grad or slots.requires_grad or any(p.requires_grad for p in self.parameters())):
            # eager training loops: forward / backward schedules replayed from CUDA graphs (graphed.py)
            g = graphed.graphs_of(self, lambda i, s: slot_attention_apply(self, i, s, True), lambda: [self._wcache])
            out, mask = g(inputs, slots)
            return out, (mask.detach() if want_mask else None)      # the segmentation mask carries no gradient (sa_diffusion.py:50-51)
        return slot_attention_apply(self, inputs, slots, want_mask)

    def forward(self, inputs, slots):
        """inputs [B, N, C], slots [B, num_slots, C] -> updated slots."""
        return self._run(inputs, slots, False)[0]

    @property
    def dtype(self):
        return self.project_k.weight.dtype

    @property
    def device(self):
        return self.project_k.weight.device


class SlotAttentionWMask(SlotAttention):
    """Also returns the last-iteration softmax-over-slots map as the segmentation mask [B, S, N]."""

    def forward(self, inputs, slots):
        return self._run(inputs, slots, True)
