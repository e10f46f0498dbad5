from . import layers, vision_transformer  # noqa: F401
from .vision_transformer import create_model  # noqa: F401
