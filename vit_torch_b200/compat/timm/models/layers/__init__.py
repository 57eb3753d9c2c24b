"""timm.models.layers symbols the reference imports: trunc_normal_, DropPath, to_2tuple."""
import collections.abc
from itertools import repeat

from vit_torch_b200.models import trunc_normal_  # noqa: F401
from vit_torch_b200.modules import DropPath  # noqa: F401


def _ntuple(n):
    def parse(x):
        if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
            return tuple(x)
        return tuple(repeat(x, n))
    return parse


to_2tuple = _ntuple(2)
