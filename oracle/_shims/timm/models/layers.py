"""Minimal stand-ins for timm.models.layers used by the reference's models/swin_transformer.py:13."""
import torch.nn as nn
from torch.nn.init import trunc_normal_  # noqa: F401


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class DropPath(nn.Identity):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        assert not drop_prob, "shim only supports drop_path == 0 (inference)"
