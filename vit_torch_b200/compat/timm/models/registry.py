"""timm.models.registry.register_model: records the reference's own constructors (used only for names the fused
package does not implement natively)."""
_registry = {}


def register_model(fn):
    _registry[fn.__name__] = fn
    return fn


def is_model(name):
    return name in _registry


def model_entrypoint(name):
    return _registry[name]
