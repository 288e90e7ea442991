"""Deterministic synthetic weights and clips.

There is no checkpoint in the reference tree (they are Google-Drive links,
README.md:61-68) and no network here, so benchmarks and parity tests run on
random weights of the right architecture.  Every tensor is drawn from a
``numpy.random.RandomState`` (the frozen legacy stream, identical on every
machine and numpy version) seeded by the CRC32 of the tensor's state-dict key,
so the golden-vector generator (which runs the real reference), the CPU oracle
and the CUDA path all see bit-identical weights without a 35 MB fixture.

Scales are chosen so that every stage does real work: LayerNorm affine terms
are non-trivial, the relative-position bias is large enough to matter in the
softmax, codebooks are spread like the reference's kaiming-normal init.
"""
from __future__ import annotations

import zlib
from typing import Dict

import numpy as np
import torch

from .spec import CodecSpec, ManifestEntry, relative_position_index


def _rs(key: str, seed: int) -> np.random.RandomState:
    return np.random.RandomState((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0xFFFFFFFF)


def _fan_in(shape) -> int:
    n = 1
    for s in shape[1:]:
        n *= s
    return max(n, 1)


def synth_tensor(entry: ManifestEntry, spec: CodecSpec, seed: int = 0) -> torch.Tensor:
    rs = _rs(entry.key, seed)
    shape = entry.shape
    role = entry.role
    if role == "window":
        return torch.hann_window(spec.win_length, periodic=True, dtype=torch.float32)
    if role == "relpos_index":
        return torch.tensor(relative_position_index(spec.window_size), dtype=torch.int64)
    if role in ("linear_w", "conv_w"):
        bound = 1.0 / np.sqrt(_fan_in(shape))
        a = rs.uniform(-bound, bound, size=shape) * np.sqrt(3.0)   # unit-gain-ish: var = 1/fan_in
    elif role == "bias":
        a = rs.uniform(-0.1, 0.1, size=shape)
    elif role == "ln_w":
        a = 1.0 + 0.1 * rs.standard_normal(size=shape)
    elif role == "ln_b":
        a = 0.1 * rs.standard_normal(size=shape)
    elif role == "relpos_table":
        a = 0.5 * rs.standard_normal(size=shape)
    elif role == "codebook":
        a = rs.standard_normal(size=shape) * np.sqrt(2.0 / shape[1])
    else:
        raise KeyError(role)
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def synth_state_dict(spec: CodecSpec, seed: int = 0) -> Dict[str, torch.Tensor]:
    return {e.key: synth_tensor(e, spec, seed) for e in spec.manifest()}


def synth_audio(batch: int, num_samples: int, seed: int = 0, scale: float = 0.1) -> torch.Tensor:
    """``scale * N(0,1)`` clips, the input BASELINE.md's CPU numbers were taken on."""
    out = np.empty((batch, num_samples), dtype=np.float32)
    for b in range(batch):
        out[b] = (_rs(f"clip:{b}", seed).standard_normal(num_samples) * scale).astype(np.float32)
    return torch.from_numpy(out)
