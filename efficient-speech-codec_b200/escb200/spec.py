"""Codec geometry derived from the reference's constructor kwargs.

Everything here is pure Python arithmetic: which Swin layers exist, what each
product quantizer looks like, and the exact ``state_dict`` key/shape manifest a
reference checkpoint carries.  The native library, the Python host module, the
test oracle and the synthetic-weight generator all read the same ``CodecSpec``
so they cannot drift apart.

Reference behaviour this mirrors (file:line under /root/reference):
  * ctor kwargs and defaults ............ esc/models/codecs.py:11-18
  * STFT parameters ..................... esc/models/base.py:22-27
  * quantizer geometry .................. esc/models/base.py:49-71,
                                          esc/modules/vq/quantization.py:21-30,380-386
  * encoder / decoder layer lists ....... esc/models/base.py:127-141, esc/models/csrvq.py:78-95
  * Swin block parameters ............... esc/modules/transformer/attention.py:119-127,181-210,258-266
  * patch (de)embed / merge / split ..... esc/modules/transformer/scale.py:26-145
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

_ESC_DEFAULTS = dict(
    in_dim=2, in_freq=192, h_dims=[45, 72, 96, 144, 192, 384], max_streams=6,
    win_len=20, hop_len=5, sr=16000, patch_size=[3, 2], swin_heads=[3, 6, 12, 24, 24],
    swin_depth=2, window_size=4, mlp_ratio=4.0, overlap=2, group_size=3,
    codebook_size=1024, codebook_dims=[8, 8, 8, 8, 8, 8], l2norm=True,
    backbone="transformer", kernel_size=[5, 2], conv_depth=1,
)
# RVQCodecs.__init__ (esc/models/codecs.py:98-119): one codebook_dim, num_rvqs residual codebooks per group
_RVQ_DEFAULTS = dict(
    in_dim=2, in_freq=192, h_dims=[45, 72, 96, 144, 192, 384], max_streams=6, backbone="transformer",
    kernel_size=[5, 2], conv_depth=1, patch_size=[3, 2], swin_heads=[3, 6, 12, 24, 24], swin_depth=2, window_size=4,
    mlp_ratio=4.0, overlap=2, num_rvqs=6, group_size=3, codebook_dim=8, codebook_size=1024, l2norm=True,
    win_len=20, hop_len=5, sr=16000,
)


@dataclass(frozen=True)
class SwinLayerSpec:
    """One reference ``TransformerLayer``: ``depth`` Swin blocks + optional resample."""
    prefix: str            # state-dict prefix, e.g. "encoder.blocks.0"
    dim: int               # channel width C of the blocks
    heads: int
    depth: int
    scale: Optional[str]   # None | "down" (PatchMerge) | "up" (PatchSplit)
    out_dim: int           # channel width after the resample (== dim when scale is None)

    @property
    def head_dim(self) -> int:
        return self.dim // self.heads


@dataclass(frozen=True)
class QuantizerSpec:
    """One reference ``ProductVectorQuantize``."""
    prefix: str
    in_dim: int            # channels C of the scale it quantizes
    in_freq: int           # frequency patches H of that scale
    overlap: int
    groups: int
    codebook_dim: int
    codebook_size: int
    vq_dims: Tuple[int, ...]

    @property
    def fix_dim(self) -> int:
        return self.in_dim * self.in_freq

    @property
    def frame_dim(self) -> int:
        return self.fix_dim * self.overlap


@dataclass(frozen=True)
class ManifestEntry:
    key: str
    shape: Tuple[int, ...]
    role: str              # linear_w | bias | ln_w | ln_b | relpos_table | relpos_index | codebook | conv_w | window
    buffer: bool = False
    dtype: str = "float32"


def split_dimension(total: int, parts: int) -> Tuple[int, ...]:
    """Near-equal split; the remainder goes to the last part (quantization.py:380-386)."""
    base = total // parts
    dims = [base] * parts
    dims[-1] = total - base * (parts - 1)
    return tuple(dims)


@dataclass
class CodecSpec:
    in_dim: int = 2
    in_freq: int = 192
    h_dims: Sequence[int] = (45, 72, 96, 144, 192, 384)
    max_streams: int = 6
    win_len: int = 20
    hop_len: int = 5
    sr: int = 16000
    patch_size: Sequence[int] = (3, 2)
    swin_heads: Sequence[int] = (3, 6, 12, 24, 24)
    swin_depth: int = 2
    window_size: int = 4
    mlp_ratio: float = 4.0
    overlap: int = 2
    group_size: int = 3
    codebook_size: int = 1024
    codebook_dims: Sequence[int] = (8, 8, 8, 8, 8, 8)
    l2norm: bool = True
    backbone: str = "transformer"
    kernel_size: Sequence[int] = (5, 2)
    conv_depth: int = 1
    extra: Dict = field(default_factory=dict)
    rvq: bool = False          # RVQCodecs (codecs.py:96-181): one ProductResidualVectorQuantize at the bottleneck
    num_rvqs: int = 0

    # ------------------------------------------------------------------ ctor
    @classmethod
    def from_kwargs(cls, **cfg) -> "CodecSpec":
        merged = dict(_ESC_DEFAULTS)
        unknown = set(cfg) - set(merged)
        if unknown:
            # same failure mode as the reference ctor: an unexpected kwarg is a TypeError
            raise TypeError(f"ESC.__init__() got an unexpected keyword argument '{sorted(unknown)[0]}'")
        merged.update(cfg)
        spec = cls(**merged)
        spec.validate()
        return spec

    @classmethod
    def from_rvq_kwargs(cls, **cfg) -> "CodecSpec":
        """Geometry of ``RVQCodecs(**cfg)``; ``codebook_dims`` is the single ``codebook_dim`` repeated."""
        merged = dict(_RVQ_DEFAULTS)
        unknown = set(cfg) - set(merged)
        if unknown:
            raise TypeError(f"RVQCodecs.__init__() got an unexpected keyword argument '{sorted(unknown)[0]}'")
        merged.update(cfg)
        d, n = merged.pop("codebook_dim"), merged.pop("num_rvqs")
        if isinstance(d, (list, tuple)):
            # configs/ablations/9kbps_rvq_conv.yaml passes a list; the reference fails inside nn.Embedding with a TypeError
            raise TypeError("empty() received an invalid combination of arguments - codebook_dim must be an int")
        spec = cls(**merged, codebook_dims=[int(d)] * merged["max_streams"], rvq=True, num_rvqs=int(n))
        spec.validate()
        if spec.num_rvqs < 1:
            raise ValueError("num_rvqs must be positive")
        return spec

    def validate(self) -> None:
        if self.backbone != "transformer":
            raise NotImplementedError(
                "esc-b200 accelerates the Swin-transformer path only; backbone='convolution' is the "
                "reference's ablation backbone and is out of scope (SURVEY.md section 2, row 10)")
        if len(self.h_dims) != self.max_streams:
            raise ValueError("len(h_dims) must equal max_streams (one scale per stream)")
        if len(self.swin_heads) != len(self.h_dims) - 1:
            raise ValueError("swin_heads needs one entry per resampling layer")
        if len(self.codebook_dims) != self.max_streams:
            raise ValueError("codebook_dims needs one entry per stream")
        if self.in_freq % self.patch_size[0] != 0:
            raise ValueError("in_freq must be a multiple of the frequency patch size")
        h = self.top_freq
        for _ in range(self.max_streams - 1):
            if h % 2:
                raise NotImplementedError("odd frequency-patch counts (PatchMerge zero-padding) are not supported")
            h //= 2
        for c, nh in zip(self.h_dims[:-1], self.swin_heads):
            if c % nh:
                raise ValueError(f"channel width {c} is not divisible by {nh} heads")
        if self.h_dims[-1] % self.swin_heads[-1]:
            raise ValueError("bottleneck width is not divisible by its head count")
        if self.window_size != 4:
            raise NotImplementedError("the native window-attention kernels are specialised for window_size=4")
        if self.in_dim != 2:
            raise ValueError("in_dim must be 2 (real/imag planes of the complex STFT)")

    # -------------------------------------------------------------- front end
    @property
    def n_fft(self) -> int:
        return (self.in_freq - 1) * 2

    @property
    def win_length(self) -> int:
        return int(self.win_len * self.sr * 1e-3)

    @property
    def hop(self) -> int:
        return int(self.hop_len * self.sr * 1e-3)

    @property
    def top_freq(self) -> int:
        """Frequency patches at the finest scale (64 for the shipped configs)."""
        return self.in_freq // self.patch_size[0]

    @property
    def bottom_freq(self) -> int:
        return self.top_freq // 2 ** (self.max_streams - 1)

    def num_frames(self, num_samples: int) -> int:
        return 1 + num_samples // self.hop

    def time_patches(self, num_samples: int) -> int:
        return self.num_frames(num_samples) // self.patch_size[1]

    def decoded_samples(self, time_patches: int) -> int:
        return self.hop * (time_patches * self.patch_size[1] - 1)

    @property
    def max_bps(self) -> float:
        # base.py:70
        return (2 / self.overlap) * self.max_streams * math.log2(self.codebook_size) * self.group_size \
            // (20 * self.patch_size[1] // 2)

    # ------------------------------------------------------------ layer lists
    @property
    def dec_h_dims(self) -> List[int]:
        return list(self.h_dims)[::-1]

    @property
    def dec_heads(self) -> List[int]:
        return list(self.swin_heads)[::-1]

    def encoder_layers(self) -> List[SwinLayerSpec]:
        h = list(self.h_dims)
        out = [SwinLayerSpec("encoder.pre_nn", h[0], self.swin_heads[0], self.swin_depth, None, h[0])]
        for i in range(len(h) - 1):
            out.append(SwinLayerSpec(f"encoder.blocks.{i}", h[i], self.swin_heads[i], self.swin_depth, "down", h[i + 1]))
        return out

    def decoder_layers(self) -> List[SwinLayerSpec]:
        h, nh = self.dec_h_dims, self.dec_heads
        out = []
        for i in range(len(h) - 1):
            out.append(SwinLayerSpec(f"decoder.blocks.{i}", h[i], nh[i], self.swin_depth, "up", h[i + 1]))
        out.append(SwinLayerSpec("decoder.post_nn", h[-1], nh[-1], self.swin_depth, None, h[-1]))
        return out

    def quantizers(self) -> List[QuantizerSpec]:
        dec = self.dec_h_dims
        H = self.top_freq
        out = []
        for i in range(self.max_streams):
            if i == 0:
                in_dim, in_freq = dec[0], H // 2 ** (self.max_streams - 1)
            else:
                in_dim, in_freq = dec[i - 1], H // 2 ** (self.max_streams - i)
            out.append(QuantizerSpec(
                prefix=f"quantizers.{i}", in_dim=in_dim, in_freq=in_freq, overlap=self.overlap,
                groups=self.group_size, codebook_dim=int(self.codebook_dims[i]),
                codebook_size=self.codebook_size,
                vq_dims=split_dimension(in_dim * in_freq * self.overlap, self.group_size)))
        return out

    # ------------------------------------------------------------- state dict
    def _swin_layer_entries(self, layer: SwinLayerSpec) -> List[ManifestEntry]:
        C, nh, ws = layer.dim, layer.heads, self.window_size
        hidden = int(C * self.mlp_ratio)
        e: List[ManifestEntry] = []
        for j in range(layer.depth):
            p = f"{layer.prefix}.swint_blocks.{j}"
            e += [
                ManifestEntry(f"{p}.norm1.weight", (C,), "ln_w"),
                ManifestEntry(f"{p}.norm1.bias", (C,), "ln_b"),
                ManifestEntry(f"{p}.attn.relative_position_bias_table", ((2 * ws - 1) ** 2, nh), "relpos_table"),
                ManifestEntry(f"{p}.attn.relative_position_index", (ws * ws, ws * ws), "relpos_index", True, "int64"),
                ManifestEntry(f"{p}.attn.qkv.weight", (3 * C, C), "linear_w"),
                ManifestEntry(f"{p}.attn.qkv.bias", (3 * C,), "bias"),
                ManifestEntry(f"{p}.attn.proj.weight", (C, C), "linear_w"),
                ManifestEntry(f"{p}.attn.proj.bias", (C,), "bias"),
                ManifestEntry(f"{p}.norm2.weight", (C,), "ln_w"),
                ManifestEntry(f"{p}.norm2.bias", (C,), "ln_b"),
                ManifestEntry(f"{p}.mlp.linear_1.weight", (hidden, C), "linear_w"),
                ManifestEntry(f"{p}.mlp.linear_1.bias", (hidden,), "bias"),
                ManifestEntry(f"{p}.mlp.linear_2.weight", (C, hidden), "linear_w"),
                ManifestEntry(f"{p}.mlp.linear_2.bias", (C,), "bias"),
            ]
        if layer.scale == "down":
            e += [
                ManifestEntry(f"{layer.prefix}.subsample.norm.weight", (2 * C,), "ln_w"),
                ManifestEntry(f"{layer.prefix}.subsample.norm.bias", (2 * C,), "ln_b"),
                ManifestEntry(f"{layer.prefix}.subsample.down.weight", (layer.out_dim, 2 * C), "linear_w"),
            ]
        elif layer.scale == "up":
            e += [
                ManifestEntry(f"{layer.prefix}.subsample.norm.weight", (C,), "ln_w"),
                ManifestEntry(f"{layer.prefix}.subsample.norm.bias", (C,), "ln_b"),
                ManifestEntry(f"{layer.prefix}.subsample.up.weight", (2 * layer.out_dim, C), "linear_w"),
            ]
        return e

    def manifest(self) -> List[ManifestEntry]:
        """Every tensor of a reference ESC (or RVQCodecs) checkpoint's ``model_state_dict``."""
        C0 = self.h_dims[0]
        pf, pt = self.patch_size
        e: List[ManifestEntry] = [
            ManifestEntry("ft.window", (self.win_length,), "window", True),
            ManifestEntry("ift.window", (self.win_length,), "window", True),
        ]
        if self.rvq:
            # ProductResidualVectorQuantize (quantization.py:276-297) of ResidualVectorQuantize (:139-168) at the bottleneck
            q0 = self.quantizers()[0]
            for m in range(q0.groups):
                p = f"quantizers.vqs.{m}"
                e.append(ManifestEntry(f"{p}.proj_down.weight", (q0.codebook_dim, q0.vq_dims[m]), "linear_w"))
                e.append(ManifestEntry(f"{p}.proj_up.weight", (q0.vq_dims[m], q0.codebook_dim), "linear_w"))
                for i in range(self.num_rvqs):
                    e.append(ManifestEntry(f"{p}.vqs.{i}.embedding.weight", (q0.codebook_size, q0.codebook_dim), "codebook"))
        for q in ([] if self.rvq else self.quantizers()):
            for g in range(q.groups):
                e.append(ManifestEntry(f"{q.prefix}.vqs.{g}.embedding.weight", (q.codebook_size, q.codebook_dim), "codebook"))
            for g in range(q.groups):
                e.append(ManifestEntry(f"{q.prefix}.down_projs.{g}.weight", (q.codebook_dim, q.vq_dims[g]), "linear_w"))
            for g in range(q.groups):
                e.append(ManifestEntry(f"{q.prefix}.up_projs.{g}.weight", (q.vq_dims[g], q.codebook_dim), "linear_w"))
        e += [
            ManifestEntry("encoder.patch_embed.proj.weight", (C0, self.in_dim, pf, pt), "conv_w"),
            ManifestEntry("encoder.patch_embed.proj.bias", (C0,), "bias"),
            ManifestEntry("encoder.patch_embed.norm.weight", (C0,), "ln_w"),
            ManifestEntry("encoder.patch_embed.norm.bias", (C0,), "ln_b"),
        ]
        for layer in self.encoder_layers():
            e += self._swin_layer_entries(layer)
        for layer in self.decoder_layers():
            e += self._swin_layer_entries(layer)
        e += [
            ManifestEntry("decoder.patch_deembed.de_proj1.weight", (C0 * pf * pt, C0, 5, 5), "conv_w"),
            ManifestEntry("decoder.patch_deembed.de_proj1.bias", (C0 * pf * pt,), "bias"),
            ManifestEntry("decoder.patch_deembed.de_proj2.weight", (self.in_dim, C0, 3, 3), "conv_w"),
            ManifestEntry("decoder.patch_deembed.de_proj2.bias", (self.in_dim,), "bias"),
        ]
        return e

    def to_kwargs(self) -> Dict:
        keys = list(_ESC_DEFAULTS)
        return {k: (list(getattr(self, k)) if isinstance(getattr(self, k), (list, tuple)) else getattr(self, k))
                for k in keys}


def relative_position_index(ws: int):
    """[ws*ws, ws*ws] table index of the Swin relative-position bias (attention.py:190-205).

    index[i, j] = (dh + ws-1) * (2*ws-1) + (dw + ws-1) with (dh, dw) the offset of token i from token j.
    """
    n = ws * ws
    idx = [[0] * n for _ in range(n)]
    for i in range(n):
        for j in range(n):
            dh = i // ws - j // ws + ws - 1
            dw = i % ws - j % ws + ws - 1
            idx[i][j] = dh * (2 * ws - 1) + dw
    return idx
