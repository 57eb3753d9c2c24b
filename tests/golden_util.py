import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unpack(d):
    return d["q"].float() * d["scale"]


def load(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def nerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
