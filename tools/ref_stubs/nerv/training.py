import torch
from torch import nn


class BaseParams:
    def get(self, key, default=None):
        return getattr(self, key, default)


class BaseModel(nn.Module):
    def load_weight(self, path, strict=True):
        ckp = torch.load(path, map_location='cpu')
        if 'state_dict' in ckp:
            ckp = ckp['state_dict']
        self.load_state_dict(ckp, strict=strict)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device


class BaseMethod:
    def __init__(self, *a, **k):
        raise NotImplementedError('stub')


class BaseDataModule:
    def __init__(self, *a, **k):
        raise NotImplementedError('stub')


class CosineAnnealingWarmupRestarts:
    def __init__(self, *a, **k):
        raise NotImplementedError('stub')
