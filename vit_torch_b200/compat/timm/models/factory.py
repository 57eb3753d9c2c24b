"""timm.models.factory.create_model(name, pretrained=False, **kwargs) as called at models/vision_all.py:186-193:
None-valued kwargs are dropped (timm behaviour), names are case-sensitive (`cait_S24_224`)."""
from .registry import _registry


def _native(name):
    from vit_torch_b200 import models as _m
    try:
        from vit_torch_b200 import cait as _c
    except Exception:  # pragma: no cover
        _c = None
    for mod in (_c, _m):
        fn = getattr(mod, name, None) if mod is not None else None
        if callable(fn) and (name.startswith("cait_") or name.startswith("deit_") or name.startswith("dino_")):
            return fn
    return None


def create_model(model_name, pretrained=False, checkpoint_path="", scriptable=None, exportable=None, no_jit=None,
                 **kwargs):
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    fn = _native(model_name)
    if fn is None:
        if model_name not in _registry:
            raise RuntimeError(f"Unknown model ({model_name})")
        fn = _registry[model_name]
    return fn(pretrained=pretrained, **kwargs)
