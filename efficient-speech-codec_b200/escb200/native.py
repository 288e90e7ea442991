"""ctypes binding of ``libescb200.so`` (the C ABI declared in ``include/escb200.h``).

This module is the only place the Python host code touches native code.  There is no
CPU fallback anywhere behind it: a missing library raises ``NativeLibraryMissing`` and a
missing CUDA device makes ``escb_create`` fail with ``ESCB_ENODEV`` (raised as
``NativeError``).  Tensors are allocated by PyTorch and passed as raw device pointers;
all work is enqueued on ``torch.cuda.current_stream()``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

ESCB_MAX_LEVELS = 8
ESCB_ABI_VERSION = 2
_LIB_NAME = "libescb200.so"

ERROR_NAMES = {0: "ESCB_OK", -1: "ESCB_EINVAL", -2: "ESCB_ENODEV", -3: "ESCB_ECUDA", -4: "ESCB_ESTATE",
               -5: "ESCB_ENOMEM", -6: "ESCB_EKEY"}

# every entry point include/escb200.h declares; tests check that the library exports all of them
EXPORTS = [
    "escb_abi_version", "escb_last_error", "escb_create", "escb_destroy", "escb_num_weights", "escb_weight_name",
    "escb_weight_numel", "escb_set_weight", "escb_finalize", "escb_time_patches", "escb_decoded_samples",
    "escb_workspace_bytes", "escb_encode", "escb_decode", "escb_forward", "escb_encode_host", "escb_decode_host",
    "escb_stft", "escb_istft", "escb_patch_embed", "escb_patch_deembed", "escb_swin_layer", "escb_pvq_encode",
    "escb_pvq_decode", "escb_codebook_argmin", "escb_launch_count", "escb_profile_begin", "escb_profile_end",
    "escb_poll_error", "escb_code_histogram", "escb_pvq_stream", "escb_forward_feat", "escb_tiling_info",
]
ESCB_NUM_OPS = 20


class NativeLibraryMissing(ImportError):
    pass


class NativeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


class EscbConfig(C.Structure):
    _fields_ = [
        ("in_freq", C.c_int32), ("win_length", C.c_int32), ("hop_length", C.c_int32),
        ("patch_freq", C.c_int32), ("patch_time", C.c_int32), ("num_levels", C.c_int32),
        ("h_dims", C.c_int32 * ESCB_MAX_LEVELS), ("swin_heads", C.c_int32 * ESCB_MAX_LEVELS),
        ("swin_depth", C.c_int32), ("window_size", C.c_int32), ("mlp_hidden_mult", C.c_int32),
        ("overlap", C.c_int32), ("group_size", C.c_int32), ("codebook_size", C.c_int32),
        ("codebook_dims", C.c_int32 * ESCB_MAX_LEVELS), ("l2norm", C.c_int32), ("num_rvqs", C.c_int32),
    ]


class EscbOpStat(C.Structure):
    _fields_ = [("name", C.c_char_p), ("launches", C.c_int64), ("ms", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double)]


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


_lib = None


def lib() -> C.CDLL:
    """Load (once) and type the library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise NativeLibraryMissing(
            f"{path} is missing: build it with `make -C efficient-speech-codec_b200/csrc` "
            f"(or `python -c 'import __graft_entry__ as g; g.build()'`).  esc-b200 has no CPU fallback.")
    L = C.CDLL(path)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    sigs = {
        "escb_abi_version": (C.c_int, []),
        "escb_last_error": (C.c_char_p, []),
        "escb_create": (C.c_int, [C.POINTER(EscbConfig), C.POINTER(vp)]),
        "escb_destroy": (None, [vp]),
        "escb_num_weights": (C.c_int, [vp]),
        "escb_weight_name": (C.c_char_p, [vp, C.c_int]),
        "escb_weight_numel": (i64, [vp, C.c_int]),
        "escb_set_weight": (C.c_int, [vp, C.c_char_p, vp, i64, C.c_int]),
        "escb_finalize": (C.c_int, [vp]),
        "escb_time_patches": (C.c_int, [vp, i64, C.POINTER(i32)]),
        "escb_decoded_samples": (i64, [vp, i32]),
        "escb_workspace_bytes": (C.c_int, [vp, i32, i32, C.POINTER(sz)]),
        "escb_encode": (C.c_int, [vp, vp, i32, i64, i32, vp, vp, sz, vp]),
        "escb_decode": (C.c_int, [vp, vp, i32, i32, i32, vp, vp, vp, sz, vp]),
        "escb_forward": (C.c_int, [vp, vp, i32, i64, i32, vp, vp, vp, vp, vp, vp, sz, vp]),
        "escb_forward_feat": (C.c_int, [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp]),
        "escb_encode_host": (C.c_int, [vp, vp, i32, i64, i32, vp, vp]),
        "escb_decode_host": (C.c_int, [vp, vp, i32, i32, i32, vp, vp]),
        "escb_stft": (C.c_int, [vp, vp, i32, i64, vp, vp, sz, vp]),
        "escb_istft": (C.c_int, [vp, vp, i32, i32, vp, vp, sz, vp]),
        "escb_patch_embed": (C.c_int, [vp, vp, i32, i32, vp, vp, sz, vp]),
        "escb_patch_deembed": (C.c_int, [vp, vp, i32, i32, vp, vp, sz, vp]),
        "escb_swin_layer": (C.c_int, [vp, i32, vp, i32, i32, i32, vp, vp, sz, vp]),
        "escb_pvq_encode": (C.c_int, [vp, i32, vp, vp, i32, i32, vp, vp, sz, vp]),
        "escb_pvq_decode": (C.c_int, [vp, i32, vp, vp, i32, i32, vp, vp, sz, vp]),
        "escb_pvq_stream": (C.c_int, [vp, i32, vp, vp, i32, i32, vp, vp, vp, sz, vp]),
        "escb_codebook_argmin": (C.c_int, [vp, i32, i32, vp, i64, vp, vp]),
        "escb_launch_count": (i64, [vp]),
        "escb_poll_error": (C.c_int, [vp]),
        "escb_tiling_info": (C.c_int, [i32, i32, i32, C.POINTER(i32)]),
        "escb_code_histogram": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, vp]),
        "escb_profile_begin": (C.c_int, [vp]),
        "escb_profile_end": (C.c_int, [vp, C.POINTER(EscbOpStat), C.POINTER(i32)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.escb_abi_version() != ESCB_ABI_VERSION:
        raise NativeLibraryMissing(f"{path}: ABI version {L.escb_abi_version()} != {ESCB_ABI_VERSION}; rebuild it")
    _lib = L
    return L


def check(code: int) -> None:
    if code != 0:
        raise NativeError(code, lib().escb_last_error().decode("utf-8", "replace"))


def make_config(spec) -> EscbConfig:
    """``CodecSpec`` -> ``escb_config`` (the ctor kwargs of esc/models/codecs.py:11-18 in C form)."""
    cfg = EscbConfig()
    cfg.in_freq = spec.in_freq
    cfg.win_length = spec.win_length
    cfg.hop_length = spec.hop
    cfg.patch_freq, cfg.patch_time = spec.patch_size
    cfg.num_levels = len(spec.h_dims)
    if cfg.num_levels > ESCB_MAX_LEVELS:
        raise ValueError(f"at most {ESCB_MAX_LEVELS} scales are supported")
    for i, v in enumerate(spec.h_dims):
        cfg.h_dims[i] = int(v)
    for i, v in enumerate(spec.swin_heads):
        cfg.swin_heads[i] = int(v)
    for i, v in enumerate(spec.codebook_dims):
        cfg.codebook_dims[i] = int(v)
    cfg.swin_depth = spec.swin_depth
    cfg.window_size = spec.window_size
    hidden = spec.mlp_ratio
    if int(hidden) != hidden:
        raise NotImplementedError("mlp_ratio must be an integer")
    cfg.mlp_hidden_mult = int(hidden)
    cfg.overlap = spec.overlap
    cfg.group_size = spec.group_size
    cfg.codebook_size = spec.codebook_size
    cfg.l2norm = 1 if spec.l2norm else 0
    cfg.num_rvqs = int(getattr(spec, "num_rvqs", 0)) if getattr(spec, "rvq", False) else 0
    return cfg


class Handle:
    """Owns one ``escb_handle``."""

    def __init__(self, spec):
        self._lib = lib()
        self._h = C.c_void_p()
        check(self._lib.escb_create(C.byref(make_config(spec)), C.byref(self._h)))

    def close(self) -> None:
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._h = None                     # before the call: close() also runs as __del__ at interpreter shutdown
            self._lib.escb_destroy(h)

    __del__ = close

    @property
    def ptr(self) -> C.c_void_p:
        return self._h

    def weight_names(self) -> List[str]:
        n = self._lib.escb_num_weights(self._h)
        return [self._lib.escb_weight_name(self._h, i).decode() for i in range(n)]

    def weight_numel(self, i: int) -> int:
        return self._lib.escb_weight_numel(self._h, i)

    def set_weight(self, name: str, tensor) -> None:
        """``tensor``: contiguous fp32 torch tensor on the CPU or on the handle's device."""
        check(self._lib.escb_set_weight(self._h, name.encode(), C.c_void_p(tensor.data_ptr()), tensor.numel(),
                                        1 if tensor.is_cuda else 0))

    def finalize(self) -> None:
        check(self._lib.escb_finalize(self._h))

    def time_patches(self, num_samples: int) -> int:
        w = C.c_int32()
        check(self._lib.escb_time_patches(self._h, num_samples, C.byref(w)))
        return w.value

    def decoded_samples(self, W: int) -> int:
        return self._lib.escb_decoded_samples(self._h, W)

    def workspace_bytes(self, batch: int, W: int) -> int:
        b = C.c_size_t()
        check(self._lib.escb_workspace_bytes(self._h, batch, W, C.byref(b)))
        return b.value

    def poll_error(self) -> None:
        """Raise if a completed kernel met an out-of-range code index since the last poll (include/escb200.h)."""
        check(self._lib.escb_poll_error(self._h))

    def launch_count(self) -> int:
        return self._lib.escb_launch_count(self._h)

    def profile_begin(self) -> None:
        check(self._lib.escb_profile_begin(self._h))

    def profile_end(self) -> dict:
        """{op name: dict(launches, ms, flops, bytes)} accumulated since profile_begin()."""
        stats = (EscbOpStat * ESCB_NUM_OPS)()
        n = C.c_int32()
        check(self._lib.escb_profile_end(self._h, stats, C.byref(n)))
        return {stats[i].name.decode(): dict(launches=stats[i].launches, ms=stats[i].ms, flops=stats[i].flops,
                                             bytes=stats[i].bytes) for i in range(n.value)}


def ptr(t: Optional["object"]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())
