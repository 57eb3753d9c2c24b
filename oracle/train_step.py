"""One optimisation step of the reference's training loop on the oracle model (TEST INFRASTRUCTURE / CPU baseline).

Follows utils_network.py:406-452 (forward :417-421, CrossEntropyLoss :429-433 created at main.py:244,
zero_grad/backward/step :439-442) with SGD momentum 0.9 (utils_network.py:120) and lr 1e-3 (main.py:81). DINO models
are built the way models/vision_all.py:154-158 builds them with pretrained=False (reset_parameters after upstream init).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import vit


def reset_parameters_like_zoo(m: nn.Module) -> None:
    for c in m.children():
        reset_parameters_like_zoo(c)
    if hasattr(m, "reset_parameters"):
        m.reset_parameters()


def build(name: str, seed: int = 0, device: str = "cpu", lr: float = 1e-3, momentum: float = 0.9):
    torch.manual_seed(seed)
    model = getattr(vit, name)(pretrained=False)
    reset_parameters_like_zoo(model)
    model = model.to(device)
    opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=momentum)
    return model, opt


def step(model, opt, x, y):
    out = model(x)
    loss = F.cross_entropy(out, y)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss.detach()
