"""`ExponentialMovingAverage` with the reference's surface (mdgen/ema.py:9-75: `params`, `decay`, `device`, `to`,
`update(model)`, `state_dict()`, `load_state_dict()`); the update `stored -= (stored - param) * (1 - decay)` of every
floating-point tensor runs in the library's `ema_update_kernel` (mdgen_ema_update) when the tensors live on the GPU."""
from __future__ import annotations

from collections import OrderedDict

import torch


class ExponentialMovingAverage:
    def __init__(self, model, decay: float):
        self.params = OrderedDict((k, v.clone().detach()) for k, v in model.state_dict().items())
        self.decay = decay
        self.device = next(model.parameters()).device
        self._model = model

    def to(self, device):
        self.params = OrderedDict((k, v.to(device)) for k, v in self.params.items())
        self.device = torch.device(device)

    def update(self, model) -> None:
        """== mdgen/ema.py:52-58 / :41-50."""
        eng = model.engine() if hasattr(model, "engine") and self.device.type == "cuda" else None
        with torch.no_grad():
            for k, v in model.state_dict().items():
                stored = self.params[k]
                if eng is not None and stored.is_cuda and stored.dtype == torch.float32 and stored.is_contiguous():
                    with torch.cuda.device(stored.device):
                        eng.ema_update(stored, v, self.decay)
                else:                                   # host copies (before .to(device)) follow the reference's arithmetic
                    diff = stored - v
                    diff *= 1 - self.decay
                    stored -= diff

    def load_state_dict(self, state_dict) -> None:
        for k in state_dict["params"].keys():
            self.params[k] = state_dict["params"][k].clone()
        self.decay = state_dict["decay"]

    def state_dict(self):
        return OrderedDict({"params": self.params, "decay": self.decay})
