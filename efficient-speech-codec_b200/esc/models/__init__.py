"""esc.models of the reference (esc/models/__init__.py:1-2): ESC and make_model, B200-native."""
from escb200.codec import ESC, RVQCodecs, make_model, model_dict  # noqa: F401
