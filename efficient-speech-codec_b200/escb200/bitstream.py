"""Compact on-disk form of the code tensor: ``ceil(log2(codebook_size))`` bits per index instead of int64.

The reference stores codes with ``torch.save`` (scripts/compress.py:35): a pickled int64 tensor, 64 bits per 10-bit
index.  At 9 kbps a 3 s clip is 6 streams x 3 groups x 150 frames = 2700 indices = 3375 bytes of payload (= 9 kbit/s),
against 21.6 KB as int64.  Layout (little endian):

    0   4   magic  b"ESCB"
    4   1   version (1)
    5   1   bits per index
    6   2   reserved (0)
    8   16  B, S, G, T as uint32
    24  ..  indices in C order of the [B, S, G, T] tensor, packed LSB-first into a byte stream

Host-side byte work on the wire format (numpy); the codec itself never sees it.
"""
from __future__ import annotations

import struct

import numpy as np
import torch

MAGIC = b"ESCB"
VERSION = 1
_HEADER = struct.Struct("<4sBBH4I")


def bits_for(codebook_size: int) -> int:
    return max(1, int(codebook_size - 1).bit_length())


def pack_codes(codes: torch.Tensor, codebook_size: int = 1024) -> bytes:
    """[B, S, G, T] integer tensor -> bytes.  Raises IndexError on an index outside [0, codebook_size)."""
    if codes.dim() != 4:
        raise ValueError("codes must have shape (Bs, num_streams, group_size, T)")
    a = codes.detach().cpu().numpy().astype(np.int64, copy=False).reshape(-1)
    if a.size and (a.min() < 0 or a.max() >= codebook_size):
        raise IndexError("index out of range in self")
    nbits = bits_for(codebook_size)
    # LSB-first: bit j of index i lands at stream position i*nbits + j
    bits = ((a[:, None] >> np.arange(nbits, dtype=np.int64)[None, :]) & 1).astype(np.uint8)
    payload = np.packbits(bits.reshape(-1), bitorder="little").tobytes()
    B, S, G, T = (int(v) for v in codes.shape)
    return _HEADER.pack(MAGIC, VERSION, nbits, 0, B, S, G, T) + payload


def unpack_codes(blob: bytes) -> torch.Tensor:
    """bytes -> [B, S, G, T] int64 tensor (what ``ESC.decode`` takes)."""
    if len(blob) < _HEADER.size:
        raise ValueError("truncated code stream")
    magic, version, nbits, _, B, S, G, T = _HEADER.unpack_from(blob, 0)
    if magic != MAGIC or version != VERSION or not 1 <= nbits <= 32:
        raise ValueError("not an ESCB code stream")
    n = B * S * G * T
    need = (n * nbits + 7) // 8
    body = np.frombuffer(blob, dtype=np.uint8, count=need, offset=_HEADER.size) if need else np.zeros(0, np.uint8)
    if body.size != need:
        raise ValueError("truncated code stream")
    bits = np.unpackbits(body, bitorder="little")[: n * nbits].reshape(n, nbits).astype(np.int64)
    vals = (bits << np.arange(nbits, dtype=np.int64)[None, :]).sum(axis=1)
    return torch.from_numpy(vals.reshape(B, S, G, T))


def save_codes(path: str, codes: torch.Tensor, codebook_size: int = 1024) -> int:
    blob = pack_codes(codes, codebook_size)
    with open(path, "wb") as f:
        f.write(blob)
    return len(blob)


def load_codes(path: str) -> torch.Tensor:
    with open(path, "rb") as f:
        return unpack_codes(f.read())
