"""``ESC`` — the reference's codec class (esc/models/codecs.py:9-94) backed by libescb200.

Same constructor kwargs, same ``state_dict`` keys/shapes, same ``encode`` / ``decode`` /
``forward`` (eval) signatures and return types, so ``scripts.compress`` / ``scripts.test``
and ``from esc import ESC`` code runs unchanged.  The module holds the checkpoint tensors as
ordinary ``nn.Parameter``s / buffers (nothing else: there are no sub-module forwards);
every computation is one C-ABI call on the current CUDA stream.  There is no CPU path:
CPU tensors are staged through pinned memory and the ``escb_*_host`` entry points, and a
machine without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import native
from .spec import CodecSpec, relative_position_index


class _Node(nn.Module):
    """Pure container; digit-named children index like an ``nn.ModuleList``."""

    def __getitem__(self, i):
        return self._modules[str(i)]

    def __len__(self):
        return sum(1 for k in self._modules if k.isdigit())

    def forward(self, *a, **k):   # pragma: no cover
        raise RuntimeError("esc-b200 modules hold checkpoint tensors only; call ESC.encode/decode/forward")


def _init_tensor(entry, spec: CodecSpec) -> torch.Tensor:
    """Same init families as the reference ctor (torch defaults; codebook.py:14; attention.py:212)."""
    shape, role = entry.shape, entry.role
    if role == "window":
        return torch.hann_window(spec.win_length, periodic=True)
    if role == "relpos_index":
        return torch.tensor(relative_position_index(spec.window_size), dtype=torch.int64)
    t = torch.empty(shape, dtype=torch.float32)
    if role in ("linear_w", "conv_w"):
        nn.init.kaiming_uniform_(t, a=math.sqrt(5))
    elif role == "bias":
        nn.init.uniform_(t, -0.05, 0.05)
    elif role == "ln_w":
        nn.init.ones_(t)
    elif role == "ln_b":
        nn.init.zeros_(t)
    elif role == "relpos_table":
        nn.init.trunc_normal_(t, std=0.02)
    elif role == "codebook":
        nn.init.kaiming_normal_(t)
    else:   # pragma: no cover
        raise KeyError(role)
    return t


class ESC(nn.Module):
    """Efficient Speech Codec, B200-native hot path."""

    def __init__(self, in_dim: int = 2, in_freq: int = 192, h_dims: list = [45, 72, 96, 144, 192, 384],
                 max_streams: int = 6, win_len: int = 20, hop_len: int = 5, sr: int = 16000,
                 patch_size: list = [3, 2], swin_heads: list = [3, 6, 12, 24, 24], swin_depth: int = 2,
                 window_size: int = 4, mlp_ratio: float = 4.,
                 overlap: int = 2, group_size: int = 3,
                 codebook_size: int = 1024, codebook_dims: list = [8, 8, 8, 8, 8, 8],
                 l2norm: bool = True, backbone: str = 'transformer',
                 kernel_size: list = [5, 2], conv_depth: int = 1) -> None:
        super().__init__()
        self.spec = CodecSpec.from_kwargs(
            in_dim=in_dim, in_freq=in_freq, h_dims=list(h_dims), max_streams=max_streams, win_len=win_len,
            hop_len=hop_len, sr=sr, patch_size=list(patch_size), swin_heads=list(swin_heads), swin_depth=swin_depth,
            window_size=window_size, mlp_ratio=mlp_ratio, overlap=overlap, group_size=group_size,
            codebook_size=codebook_size, codebook_dims=list(codebook_dims), l2norm=l2norm, backbone=backbone,
            kernel_size=list(kernel_size), conv_depth=conv_depth)
        self._init_common(in_freq, in_dim, max_streams, h_dims)

    def _init_common(self, in_freq, in_dim, max_streams, h_dims) -> None:
        # attributes the reference exposes (base.py:16-20, 70)
        self.in_freq, self.in_dim = in_freq, in_dim
        self.max_streams = max_streams
        self.enc_h_dims = list(h_dims)
        self.dec_h_dims = list(h_dims)[::-1]
        self.max_bps = self.spec.max_bps
        for e in self.spec.manifest():
            self._register(e.key, _init_tensor(e, self.spec), e.buffer)
        self._handles: Dict[torch.device, native.Handle] = {}
        self._synced: Dict[torch.device, tuple] = {}
        self._tensors: Optional[list] = None         # [(checkpoint key, tensor)] in the native manifest's order
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.refresh_weights())
        self._workspace: Dict[torch.device, torch.Tensor] = {}
        self._pinned: Dict[str, torch.Tensor] = {}

    # ------------------------------------------------------------------ module tree
    def _register(self, key: str, tensor: torch.Tensor, is_buffer: bool) -> None:
        *path, leaf = key.split(".")
        node: nn.Module = self
        for name in path:
            if name not in node._modules:
                node.add_module(name, _Node())
            node = node._modules[name]
        if is_buffer:
            node.register_buffer(leaf, tensor)
        else:
            node.register_parameter(leaf, nn.Parameter(tensor))

    # ------------------------------------------------------------------ native plumbing
    def _exec_device(self, t: torch.Tensor) -> torch.device:
        if t.is_cuda:
            return t.device
        if not torch.cuda.is_available():
            raise RuntimeError("esc-b200 has no CPU fallback: a CUDA device (B200, sm_100a) is required")
        return torch.device("cuda", torch.cuda.current_device())

    def refresh_weights(self) -> None:
        """Force a re-pack of the native weight images on the next call.

        ``_handle`` notices re-assigned tensors (``.to()``, ``load_state_dict``) and in-place autograd-visible updates
        through ``(data_ptr, _version)``; writes that bypass the version counter (``p.data.copy_()``, EMA weight
        surgery) are invisible to it - call this after such a write."""
        self._tensors = None
        self._synced.clear()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)         # .to() / .cuda() / .float(): the tensors are replaced
        self._tensors = None
        return out

    def _handle(self, dev: torch.device) -> native.Handle:
        """Handle for ``dev`` with the module's current tensors loaded (re-packed only when they changed)."""
        h = self._handles.get(dev)
        if h is None:
            with torch.cuda.device(dev):
                h = native.Handle(self.spec)
            self._handles[dev] = h
        if self._tensors is None:                            # one state_dict() walk per weight change, not per call
            sd = self.state_dict(keep_vars=True)
            self._tensors = [(name, sd[name]) for name in h.weight_names()]
        stamp = tuple((v.data_ptr(), v._version) for _, v in self._tensors)
        if self._synced.get(dev) != stamp:
            with torch.cuda.device(dev):
                for name, v in self._tensors:
                    t = v.detach()
                    if t.dtype != torch.float32 or not t.is_contiguous():
                        t = t.float().contiguous()
                    if t.is_cuda and t.device != dev:
                        t = t.cpu()
                    h.set_weight(name, t)
                h.finalize()
            self._synced[dev] = stamp
        return h

    def _ws(self, dev: torch.device, nbytes: int) -> torch.Tensor:
        ws = self._workspace.get(dev)
        if ws is None or ws.numel() < nbytes:
            self._workspace[dev] = ws = None   # release before growing
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._workspace[dev] = ws
        return ws

    def _pin(self, name: str, shape, dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        buf = self._pinned.get(name)
        if buf is None or buf.dtype != dtype or buf.numel() < n:
            buf = torch.empty(max(n, 1), dtype=dtype).pin_memory()
            self._pinned[name] = buf
        return buf[:n].view(*shape)

    def _host_in(self, name: str, t: torch.Tensor, dtype) -> torch.Tensor:
        """A host tensor the DMA engine can read: ``t`` itself when it is already pinned, contiguous and of the right
        dtype, else a copy in a cached pinned staging buffer."""
        if t.dtype == dtype and t.is_contiguous() and t.is_pinned():
            return t
        buf = self._pin(name, tuple(t.shape), dtype)
        buf.copy_(t)
        return buf

    @staticmethod
    def _stream(dev: torch.device) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def time_patches(self, num_samples: int) -> int:
        """W for a clip length; raises like the reference's assert (quantization.py:407) on a bad length."""
        W = self.spec.time_patches(num_samples)
        if W <= 0 or W % self.spec.overlap:
            raise AssertionError("Time dimension must be multiple of overlap")
        return W

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def encode(self, x: torch.Tensor, num_streams: int = 6):
        """audio [Bs, L] -> (codes [Bs, num_streams, group_size, W/overlap] int64, (H, W)) — codecs.py:68-81."""
        if x.dim() != 2:
            raise ValueError("x must have shape (Bs, L)")
        B, Ls = x.shape
        W = self.time_patches(Ls)
        dev = self._exec_device(x)
        h = self._handle(dev)
        lib = native.lib()
        shape = (B, num_streams, self.spec.group_size, W // self.spec.overlap)
        with torch.cuda.device(dev):
            if x.is_cuda:
                xx = x.contiguous().float()
                codes = torch.empty(shape, dtype=torch.int64, device=dev)
                nbytes = h.workspace_bytes(B, W)
                ws = self._ws(dev, nbytes)
                native.check(lib.escb_encode(h.ptr, native.ptr(xx), B, Ls, num_streams, native.ptr(codes),
                                             native.ptr(ws), ws.numel(), self._stream(dev)))
            else:
                xin = self._host_in("audio_in", x, torch.float32)
                codes = torch.empty(shape, dtype=torch.int64, pin_memory=True)    # caching host allocator: no clone
                native.check(lib.escb_encode_host(h.ptr, native.ptr(xin), B, Ls, num_streams, native.ptr(codes),
                                                  self._stream(dev)))
        return codes, (self.spec.bottom_freq, W)

    @torch.no_grad()
    def decode(self, codes: torch.Tensor, feat_shape: Tuple[int, int] = (2, 1000)):
        """codes -> audio [Bs, hop*(patch_time*W - 1)] — codecs.py:83-94."""
        if codes.dim() != 4 or codes.shape[2] != self.spec.group_size:
            raise ValueError("codes must have shape (Bs, num_streams, group_size, T)")
        B, S, _, T = codes.shape
        W = int(feat_shape[1])
        if T * self.spec.overlap != W:
            raise ValueError(f"codes hold {T} frames but feat_shape says W={W}")
        dev = self._exec_device(codes)
        h = self._handle(dev)
        lib = native.lib()
        n_out = h.decoded_samples(W)
        with torch.cuda.device(dev):
            if codes.is_cuda:
                cc = codes.contiguous().to(torch.int64)
                audio = torch.empty((B, n_out), dtype=torch.float32, device=dev)
                ws = self._ws(dev, h.workspace_bytes(B, W))
                native.check(lib.escb_decode(h.ptr, native.ptr(cc), B, S, W, native.ptr(audio), None, native.ptr(ws),
                                             ws.numel(), self._stream(dev)))
            else:
                # host codes come from saved encoded_*.pth files: same failure as F.embedding (codebook.py:53)
                if codes.numel() and (int(codes.min()) < 0 or int(codes.max()) >= self.spec.codebook_size):
                    raise IndexError("index out of range in self")
                cin = self._host_in("codes_in", codes, torch.int64)
                audio = torch.empty((B, n_out), dtype=torch.float32, pin_memory=True)
                native.check(lib.escb_decode_host(h.ptr, native.ptr(cin), B, S, W, native.ptr(audio), self._stream(dev)))
        return audio

    def forward_one_step(self, x, x_feat=None, num_streams=6, freeze_codebook=False):
        """Eval-mode fused encode+decode — codecs.py:30-46, csrvq.py:97-129."""
        if self.training:
            raise RuntimeError("esc-b200 implements the inference path only: call model.eval() first "
                               "(training stays with the reference, SURVEY.md section 8f)")
        if x_feat is not None:
            return self._forward_from_feat(x, x_feat, int(num_streams))
        if x.dim() != 2:
            raise ValueError("x must have shape (Bs, L)")
        B, Ls = x.shape
        W = self.time_patches(Ls)
        S = int(num_streams)
        dev = self._exec_device(x)
        h = self._handle(dev)
        lib = native.lib()
        T = self.spec.num_frames(Ls)
        F = self.spec.in_freq
        with torch.cuda.device(dev), torch.no_grad():
            xx = x.detach().to(dev, non_blocking=True).contiguous().float()
            codes = torch.empty((B, S, self.spec.group_size, W // self.spec.overlap), dtype=torch.int64, device=dev)
            audio = torch.empty((B, h.decoded_samples(W)), dtype=torch.float32, device=dev)
            raw = torch.empty((B, 2, F, T), dtype=torch.float32, device=dev)
            rec = torch.empty((B, 2, F, self.spec.patch_size[1] * W), dtype=torch.float32, device=dev)
            loss = torch.empty((B,), dtype=torch.float32, device=dev)
            ws = self._ws(dev, h.workspace_bytes(B, W))
            native.check(lib.escb_forward(h.ptr, native.ptr(xx), B, Ls, S, native.ptr(codes), native.ptr(audio),
                                          native.ptr(raw), native.ptr(rec), native.ptr(loss), native.ptr(ws),
                                          ws.numel(), self._stream(dev)))
        out = {"cm_loss": loss, "cb_loss": loss.clone(), "raw_audio": x, "recon_audio": audio, "raw_feat": raw,
               "recon_feat": rec, "codes": codes}
        if not x.is_cuda:
            out = {k: (v.cpu() if k != "raw_audio" else v) for k, v in out.items()}
        return out

    def _forward_from_feat(self, x, x_feat, S):
        """``forward`` with a precomputed complex STFT ``x_feat`` [Bs, F, T, 2] (codecs.py:33-34): no STFT is run."""
        if x_feat.dim() != 4 or x_feat.shape[1] != self.spec.in_freq or x_feat.shape[3] != 2:
            raise ValueError("x_feat must have shape (Bs, in_freq, T, 2)")
        B, F, T, _ = x_feat.shape
        pt = self.spec.patch_size[1]
        W = (T - pt) // pt + 1
        if W <= 0 or W % self.spec.overlap:
            raise AssertionError("Time dimension must be multiple of overlap")
        dev = self._exec_device(x_feat)
        h = self._handle(dev)
        with torch.cuda.device(dev), torch.no_grad():
            planes = x_feat.detach().to(dev).permute(0, 3, 1, 2).contiguous().float()          # "b h w c -> b c h w"
            codes = torch.empty((B, S, self.spec.group_size, W // self.spec.overlap), dtype=torch.int64, device=dev)
            audio = torch.empty((B, h.decoded_samples(W)), dtype=torch.float32, device=dev)
            rec = torch.empty((B, 2, F, pt * W), dtype=torch.float32, device=dev)
            loss = torch.empty((B,), dtype=torch.float32, device=dev)
            ws = self._ws(dev, h.workspace_bytes(B, W))
            native.check(native.lib().escb_forward_feat(h.ptr, native.ptr(planes), B, T, S, native.ptr(codes), native.ptr(audio),
                                                        native.ptr(rec), native.ptr(loss), native.ptr(ws), ws.numel(),
                                                        self._stream(dev)))
        out = {"cm_loss": loss, "cb_loss": loss.clone(), "raw_audio": x, "recon_audio": audio, "raw_feat": planes,
               "recon_feat": rec, "codes": codes}
        if not x_feat.is_cuda:
            out = {k: (v.cpu() if k != "raw_audio" else v) for k, v in out.items()}
        return out

    def forward(self, x, x_feat, num_streams, freeze_codebook=False):
        num_streams = self.max_streams if freeze_codebook else num_streams
        return self.forward_one_step(x, x_feat, num_streams, freeze_codebook)

    # the reference's two front-end helpers (base.py:29-47), exposed for parity tests and callers that use them
    @torch.no_grad()
    def spec_transform(self, x: torch.Tensor) -> torch.Tensor:
        B, Ls = x.shape
        dev = self._exec_device(x)
        h = self._handle(dev)
        with torch.cuda.device(dev):
            xx = x.to(dev).contiguous().float()
            T = self.spec.num_frames(Ls)
            out = torch.empty((B, 2, self.spec.in_freq, T), dtype=torch.float32, device=dev)
            ws = self._ws(dev, h.workspace_bytes(B, max(self.spec.overlap, 2 * ((T + 3) // 4))))
            native.check(native.lib().escb_stft(h.ptr, native.ptr(xx), B, Ls, native.ptr(out), native.ptr(ws),
                                                ws.numel(), self._stream(dev)))
        return out if x.is_cuda else out.cpu()

    @torch.no_grad()
    def audio_reconstruct(self, feat: torch.Tensor) -> torch.Tensor:
        B, _, F, T = feat.shape
        dev = self._exec_device(feat)
        h = self._handle(dev)
        with torch.cuda.device(dev):
            ff = feat.to(dev).contiguous().float()
            out = torch.empty((B, self.spec.hop * (T - 1)), dtype=torch.float32, device=dev)
            ws = self._ws(dev, h.workspace_bytes(B, max(self.spec.overlap, 2 * ((T + 3) // 4))))
            native.check(native.lib().escb_istft(h.ptr, native.ptr(ff), B, T, native.ptr(out), native.ptr(ws),
                                                 ws.numel(), self._stream(dev)))
        return out if feat.is_cuda else out.cpu()


class RVQCodecs(ESC):
    """The reference's RVQ ablation codec (esc/models/codecs.py:96-181): the same Swin encoder, ONE
    ``ProductResidualVectorQuantize`` at the bottleneck (``num_rvqs`` residual codebooks per group in the projected
    space, esc/modules/vq/quantization.py:139-378) and the plain up-sampling ``Decoder`` (esc/models/base.py:161-203).
    Same ``encode`` / ``decode`` / ``forward`` (eval) surface and code tensor shape ``[Bs, num_streams, group_size, T]``;
    the native library runs it through the kernels of the ESC path plus a residual-chain kernel."""

    def __init__(self, in_dim: int = 2, in_freq: int = 192, h_dims: list = [45, 72, 96, 144, 192, 384], max_streams: int = 6,
                 backbone: str = 'transformer', kernel_size: list = [5, 2], conv_depth: int = 1, patch_size: list = [3, 2],
                 swin_heads: list = [3, 6, 12, 24, 24], swin_depth: int = 2, window_size: int = 4, mlp_ratio: float = 4.,
                 overlap: int = 2, num_rvqs: int = 6, group_size: int = 3, codebook_dim: int = 8, codebook_size: int = 1024,
                 l2norm: bool = True, win_len: int = 20, hop_len: int = 5, sr: int = 16000) -> None:
        nn.Module.__init__(self)
        self.spec = CodecSpec.from_rvq_kwargs(
            in_dim=in_dim, in_freq=in_freq, h_dims=list(h_dims), max_streams=max_streams, backbone=backbone,
            kernel_size=list(kernel_size), conv_depth=conv_depth, patch_size=list(patch_size), swin_heads=list(swin_heads),
            swin_depth=swin_depth, window_size=window_size, mlp_ratio=mlp_ratio, overlap=overlap, num_rvqs=num_rvqs,
            group_size=group_size, codebook_dim=codebook_dim, codebook_size=codebook_size, l2norm=l2norm, win_len=win_len,
            hop_len=hop_len, sr=sr)
        self._init_common(in_freq, in_dim, max_streams, h_dims)
        self.dims = 3


model_dict = {"csvq+swinT": ESC, "csvq+conv": ESC, "rvq+swinT": RVQCodecs, "rvq+conv": RVQCodecs}


def make_model(model_config, model_name: str = "csvq+swinT"):
    """codecs.py:190-200.  ``model_name`` defaults to ``csvq+swinT`` so the reference's own one-argument call in
    scripts/compress.py:22 works; the ``rvq+*`` families are the reference's ablation baselines."""
    if model_name not in model_dict:
        raise NotImplementedError(f"{model_name} is not valid within [csvq+conv, csvq+swinT, rvq+conv, rvq+swinT]")
    cfg = model_config if isinstance(model_config, dict) else vars(model_config)
    return model_dict[model_name](**cfg)
