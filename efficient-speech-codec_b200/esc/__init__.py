"""Drop-in import path of the reference package (`from esc import ESC`, esc/__init__.py:1), backed by libescb200."""
from .models import ESC, RVQCodecs, make_model  # noqa: F401


def __getattr__(name):
    if name in ("Discriminator",):
        raise NotImplementedError(f"esc.{name} is outside the accelerated hot path (SURVEY.md section 8f); "
                                  "use the reference package for it")
    raise AttributeError(name)
