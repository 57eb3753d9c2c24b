from . import factory, layers, registry, vision_transformer  # noqa: F401
from .factory import create_model  # noqa: F401
from .registry import register_model  # noqa: F401
