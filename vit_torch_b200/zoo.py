"""Mirror of the reference's classifier-head factory (models/vision_all.py:299-320) on the fused kernels.

`get_classifier_head(in_features, classifier_units, classifier_act)` returns a module with the same structure and
state_dict keys as the reference's `nn.Sequential(Linear, GELU, ..., Linear(bias=False))` ("0.weight", "0.bias",
"2.weight", ...) whose forward runs vit_torch_b200.functional.HeadFn. `patch_reference_zoo(VisionModelZoo)` swaps the
factory inside the reference's own zoo class so that `main.py` fine-tune / --lineareval train the fused head."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as Fn


class ClassifierHead(nn.Sequential):
    """Sequential(Linear(bias=True), GELU, ..., Linear(bias=False)); only exact-erf GELU is fused."""

    def forward(self, x):
        linears = [m for m in self if isinstance(m, nn.Linear)]
        acts = [m for m in self if not isinstance(m, nn.Linear)]
        if not linears:
            return x
        if not x.is_cuda or any(not isinstance(a, nn.GELU) or getattr(a, "approximate", "none") != "none" for a in acts):
            raise NotImplementedError("fused classifier head needs CUDA input and exact GELU activations")
        args = []
        for lin in linears:
            args += [lin.weight, lin.bias]
        lead = x.shape[:-1]
        y = Fn.HeadFn.apply(x.reshape(-1, x.shape[-1]).float(), len(linears), *args)
        return y.reshape(*lead, y.shape[-1])


def get_classifier_head(in_features, classifier_units=None, classifier_act=None):
    """Same signature and layer layout as VisionModelZoo.get_classifier_head: the last Linear has no bias and no
    activation; one shared activation instance follows every other Linear."""
    classifier_act = classifier_act if classifier_act is not None else nn.GELU()
    layers = []
    if isinstance(classifier_units, int):
        classifier_units = [classifier_units]
    if isinstance(classifier_units, list):
        for i, v in enumerate(classifier_units):
            fin = in_features if i == 0 else classifier_units[i - 1]
            last = i >= len(classifier_units) - 1
            layers.append(nn.Linear(in_features=fin, out_features=v, bias=not last))
            if not last:
                layers.append(classifier_act)
    return ClassifierHead(*layers)


def patch_reference_zoo(zoo_cls):
    """Make the reference's VisionModelZoo build fused heads (classmethod swap; everything else untouched)."""
    zoo_cls.get_classifier_head = classmethod(lambda cls, in_features, classifier_units=None, classifier_act=nn.GELU():
                                              get_classifier_head(in_features, classifier_units, classifier_act))
    return zoo_cls
