"""The reference's optimisation step (utils_network.py:406-452) around the fused model: forward, CrossEntropyLoss,
zero_grad / backward, (DP: bucketed all-reduce), SGD with momentum 0.9 (utils_network.py:119-126)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def reset_parameters_like_zoo(m: nn.Module) -> None:
    """VisionModelZoo.reset_parameters (models/vision_all.py:322-329): recursive reset_parameters(); applied to DINO
    models built with pretrained=False (models/vision_all.py:157-158)."""
    for c in m.children():
        reset_parameters_like_zoo(c)
    if hasattr(m, "reset_parameters"):
        m.reset_parameters()


class Trainer:
    def __init__(self, model: nn.Module, lr: float = 1e-3, momentum: float = 0.9, reducer=None):
        self.model = model
        self.reducer = reducer
        self.opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=lr, momentum=momentum)
        self.extra_launches_per_step = 0

    def step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        out = self.model(x)
        loss = F.cross_entropy(out, y)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.step()
        return loss.detach()
