"""Stand-in for pytorch_lightning (reference: mdgen/wrapper.py:7,46-50). No arithmetic lives
in Lightning; the reference only needs an nn.Module base with `save_hyperparameters` and a
`device` property."""
import torch
import torch.nn as nn
from . import callbacks  # noqa: F401


class LightningModule(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.trainer = None
        self.current_epoch = 0

    def save_hyperparameters(self, *a, **k):
        return None

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")


class Trainer:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError("pytorch_lightning stand-in: Trainer is not available")
