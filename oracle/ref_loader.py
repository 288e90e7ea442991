"""Stage and import the REAL reference (yzGuu830/efficient-speech-codec) as a measurement arm.

TEST / BENCH INFRASTRUCTURE ONLY - nothing under efficient-speech-codec_b200/ imports this module.

The reference is pure Python, so it can be run but not "built".  ``stage()`` copies the three
directories its hot path needs - ``esc/``, ``scripts/``, ``configs/`` - from ``/root/reference`` (read-only, present
in the build container only) into ``baseline/_ref/``, which is git-ignored (the history stays free of reference
sources) but NOT gpurun-ignored, so it travels to the GPU box with the snapshot.  ``__graft_entry__.build()`` calls
it whenever ``/root/reference`` exists.

``load_esc()`` imports the staged package under the alias ``esc_ref`` (the reference uses relative imports inside
``esc/``; the alias keeps it apart from this repository's own drop-in ``esc`` package) after installing shims for the
two third-party imports the reference makes at module scope and that are absent from the image:
``timm.models.layers.{trunc_normal_, to_2tuple}`` (esc/modules/transformer/attention.py:6) and
``audiotools.{AudioSignal, STFTParams, ml.BaseModel}`` (esc/models/discriminator.py:8-10).
"""
import collections.abc
import importlib.util
import itertools
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference"
STAGED = os.path.join(ROOT, "baseline", "_ref")
_PARTS = ("esc", "scripts", "configs")


def stage(src: str = REF_SRC, dst: str = STAGED) -> bool:
    """Copy the reference's hot-path directories to ``dst``; returns False when ``src`` does not exist."""
    if not os.path.isdir(os.path.join(src, "esc")):
        return False
    os.makedirs(dst, exist_ok=True)
    for part in _PARTS:
        d = os.path.join(dst, part)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(src, part), d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in ("LICENSE",):
        if os.path.exists(os.path.join(src, f)):
            shutil.copy(os.path.join(src, f), os.path.join(dst, f))
    return True


def reference_root():
    """Directory holding the reference's ``esc`` package: the staged copy, else the container's read-only checkout."""
    for cand in (STAGED, REF_SRC):
        if os.path.isfile(os.path.join(cand, "esc", "__init__.py")):
            return cand
    return None


def install_shims() -> None:
    import torch
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.trunc_normal_ = torch.nn.init.trunc_normal_

        def to_2tuple(x):
            if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
                return tuple(x)
            return tuple(itertools.repeat(x, 2))
        layers.to_2tuple = to_2tuple
        timm.models, models.layers = models, layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    if "audiotools" not in sys.modules:
        at = types.ModuleType("audiotools")
        at.AudioSignal = type("AudioSignal", (), {})
        at.STFTParams = type("STFTParams", (), {})
        ml = types.ModuleType("audiotools.ml")
        ml.BaseModel = torch.nn.Module
        at.ml = ml
        sys.modules.update({"audiotools": at, "audiotools.ml": ml})


def load_esc():
    """The reference's ``esc`` package as module ``esc_ref`` (None when it is not available on this machine)."""
    if "esc_ref" in sys.modules:
        return sys.modules["esc_ref"]
    root = reference_root()
    if root is None:
        return None
    install_shims()
    pkg_dir = os.path.join(root, "esc")
    spec = importlib.util.spec_from_file_location("esc_ref", os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["esc_ref"] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        del sys.modules["esc_ref"]
        raise
    mod.__reference_root__ = root
    return mod


def make_reference_model(cfg: dict, state_dict: dict, model_name: str = "csvq+swinT"):
    """``make_model(cfg, model_name)`` of the reference (esc/models/codecs.py:190-200) with ``state_dict`` loaded
    strictly, in eval mode; None when the reference is unavailable."""
    mod = load_esc()
    if mod is None:
        return None
    from esc_ref.models import make_model
    model = make_model(dict(cfg), model_name)
    model.load_state_dict(state_dict, strict=True)
    return model.eval()


if __name__ == "__main__":
    ok = stage()
    print(f"staged {REF_SRC} -> {STAGED}" if ok else f"{REF_SRC} not present; nothing staged")
