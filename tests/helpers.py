"""Shared fixtures for the parity tests: the reference's three shipped model configs
(/root/reference/configs/9kbps_esc_{base,base_adv,large}.yaml, `model:` block) and builders."""
import numpy as np
import torch

BASE = dict(backbone="transformer", in_dim=2, in_freq=192, h_dims=[45, 72, 96, 144, 192, 384], max_streams=6,
            win_len=20, hop_len=5, sr=16000, patch_size=[3, 2], swin_heads=[3, 6, 12, 24, 24], swin_depth=2,
            window_size=4, mlp_ratio=4.0, overlap=2, group_size=3, codebook_size=1024,
            codebook_dims=[32, 32, 16, 12, 8, 6], l2norm=True)
ADV = dict(BASE, codebook_dims=[8] * 6)
LARGE = dict(BASE, swin_depth=4, codebook_dims=[8] * 6)


def make_oracle(cfg, seed):
    from escb200.spec import CodecSpec
    from escb200.synthetic import synth_state_dict
    from oracle.esc_oracle import EscOracle
    sd = synth_state_dict(CodecSpec.from_kwargs(**cfg), seed)
    return EscOracle(cfg, sd), sd


def i64(a):
    return torch.from_numpy(np.asarray(a).astype(np.int64))


def make_native(cfg, seed, device="cuda"):
    """The product: escb200.codec.ESC with the same synthetic checkpoint the oracle got."""
    from escb200.codec import ESC
    from escb200.spec import CodecSpec
    from escb200.synthetic import synth_state_dict
    sd = synth_state_dict(CodecSpec.from_kwargs(**cfg), seed)
    m = ESC(**cfg)
    m.load_state_dict(sd, strict=True)
    return m.eval().to(device), sd


class Unit:
    """Thin caller of the per-module C-ABI entry points (include/escb200.h, "unit entry points")."""

    def __init__(self, model):
        import torch
        from escb200 import native
        self.m, self.native, self.torch = model, native, torch
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.h = model._handle(self.dev)
        self.lib = native.lib()

    def _ws(self, B, W):
        return self.m._ws(self.dev, self.h.workspace_bytes(B, max(2, W + (W % 2))))

    def _st(self):
        return self.m._stream(self.dev)

    def swin_layer(self, li, x, H, W, out_shape):
        t, n = self.torch, self.native
        x = x.to(self.dev).contiguous()
        y = t.empty(out_shape, dtype=t.float32, device=self.dev)
        ws = self._ws(x.shape[0], W)
        n.check(self.lib.escb_swin_layer(self.h.ptr, li, n.ptr(x), x.shape[0], H, W, n.ptr(y), n.ptr(ws), ws.numel(), self._st()))
        return y.cpu()

    def patch_embed(self, planes):
        t, n = self.torch, self.native
        B, _, F, T = planes.shape
        W = (T - 2) // 2 + 1
        p = planes.to(self.dev).contiguous()
        y = t.empty((B, self.m.spec.top_freq * W, self.m.spec.h_dims[0]), dtype=t.float32, device=self.dev)
        ws = self._ws(B, W + 2)
        n.check(self.lib.escb_patch_embed(self.h.ptr, n.ptr(p), B, T, n.ptr(y), n.ptr(ws), ws.numel(), self._st()))
        return y.cpu()

    def patch_deembed(self, tok, W):
        t, n = self.torch, self.native
        B = tok.shape[0]
        x = tok.to(self.dev).contiguous()
        y = t.empty((B, 2, self.m.spec.in_freq, 2 * W), dtype=t.float32, device=self.dev)
        ws = self._ws(B, W)
        n.check(self.lib.escb_patch_deembed(self.h.ptr, n.ptr(x), B, W, n.ptr(y), n.ptr(ws), ws.numel(), self._st()))
        return y.cpu()

    def pvq_encode(self, q, enc, dec, W):
        t, n = self.torch, self.native
        B = enc.shape[0]
        e = enc.to(self.dev).contiguous()
        d = None if dec is None else dec.to(self.dev).contiguous()
        codes = t.empty((B, 3, W // 2), dtype=t.int64, device=self.dev)
        ws = self._ws(B, W)
        n.check(self.lib.escb_pvq_encode(self.h.ptr, q, n.ptr(e), n.ptr(d), B, W, n.ptr(codes), n.ptr(ws), ws.numel(), self._st()))
        return codes.cpu()

    def pvq_decode(self, q, codes, dec, W, shape):
        t, n = self.torch, self.native
        B = codes.shape[0]
        c = codes.to(self.dev).contiguous()
        d = None if dec is None else dec.to(self.dev).contiguous()
        out = t.empty(shape, dtype=t.float32, device=self.dev)
        ws = self._ws(B, W)
        n.check(self.lib.escb_pvq_decode(self.h.ptr, q, n.ptr(c), n.ptr(d), B, W, n.ptr(out), n.ptr(ws), ws.numel(), self._st()))
        return out.cpu()

    def pvq_stream(self, q, enc, dec, W, refine=True):
        """escb_pvq_stream: codes = vq.encode(enc - dec) and out = vq.decode(codes) + dec in one launch."""
        t, n = self.torch, self.native
        B = enc.shape[0]
        e = enc.to(self.dev).contiguous()
        d = None if dec is None else dec.to(self.dev).contiguous()
        codes = t.empty((B, 3, W // 2), dtype=t.int64, device=self.dev)
        out = t.empty_like(e) if refine else None
        ws = self._ws(B, W)
        n.check(self.lib.escb_pvq_stream(self.h.ptr, q, n.ptr(e), n.ptr(d), B, W, n.ptr(codes), n.ptr(out), n.ptr(ws), ws.numel(), self._st()))
        return codes.cpu(), (None if out is None else out.cpu())

    def argmin(self, q, g, z):
        t, n = self.torch, self.native
        zz = z.to(self.dev).contiguous()
        rows = zz.shape[0]
        idx = t.empty((rows,), dtype=t.int64, device=self.dev)
        n.check(self.lib.escb_codebook_argmin(self.h.ptr, q, g, n.ptr(zz), rows, n.ptr(idx), self._st()))
        return idx.cpu()
