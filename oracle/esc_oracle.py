"""CPU oracle for the ESC encode/decode hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The shipped path
(``efficient-speech-codec_b200/esc``) must never call into it and fails loudly
when the CUDA library is missing.

It is a functional, module-free restatement of the reference's algorithm in
fp32 ATen CPU ops, operating on a flat ``state_dict`` with the reference's own
key names.  Parity status: **pinned** — ``tests/golden/make_golden.py`` runs the
real reference (imported read-only from /root/reference) on seeded weights and
inputs and commits its outputs (codes, audio, per-stage taps) under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against
them.  The reference itself ships no tests or golden vectors.

Each function cites the reference lines it follows (paths under /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- geometry
def _split_dimension(total: int, parts: int) -> List[int]:
    # esc/modules/vq/quantization.py:380-386
    base = total // parts
    return [base] * (parts - 1) + [total - base * (parts - 1)]


class OracleConfig:
    """The subset of ``ESC.__init__`` kwargs (codecs.py:11-18) the hot path depends on."""

    def __init__(self, **kw):
        d = dict(in_dim=2, in_freq=192, h_dims=[45, 72, 96, 144, 192, 384], max_streams=6, win_len=20,
                 hop_len=5, sr=16000, patch_size=[3, 2], swin_heads=[3, 6, 12, 24, 24], swin_depth=2,
                 window_size=4, mlp_ratio=4.0, overlap=2, group_size=3, codebook_size=1024,
                 codebook_dims=[8] * 6, l2norm=True, backbone="transformer", kernel_size=[5, 2], conv_depth=1)
        d.update(kw)
        self.__dict__.update(d)
        self.n_fft = (self.in_freq - 1) * 2                      # base.py:22
        self.win_length = int(self.win_len * self.sr * 1e-3)     # base.py:23
        self.hop = int(self.hop_len * self.sr * 1e-3)            # base.py:24
        self.top_freq = self.in_freq // self.patch_size[0]
        self.dec_h_dims = list(self.h_dims)[::-1]
        self.dec_heads = list(self.swin_heads)[::-1]

    def quantizer_geometry(self, i: int) -> Tuple[int, int]:
        """(in_dim, in_freq) of stream ``i`` — base.py:55-68."""
        if i == 0:
            return self.dec_h_dims[0], self.top_freq // 2 ** (self.max_streams - 1)
        return self.dec_h_dims[i - 1], self.top_freq // 2 ** (self.max_streams - i)


# --------------------------------------------------------------------------- STFT front / back end
def stft_planes(x: Tensor, window: Tensor, cfg: OracleConfig) -> Tensor:
    """audio [B, L] -> re/im planes [B, 2, F, T].

    base.py:29-37 via torchaudio ``Spectrogram(power=None)`` -> ``torch.stft(center=True,
    pad_mode='reflect', normalized=False, onesided=True)`` with a periodic hann window of
    ``win_length`` centred inside ``n_fft``.
    """
    spec = torch.stft(x, cfg.n_fft, hop_length=cfg.hop, win_length=cfg.win_length, window=window,
                      center=True, pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    return torch.view_as_real(spec).permute(0, 3, 1, 2)


def istft_audio(planes: Tensor, window: Tensor, cfg: OracleConfig) -> Tensor:
    """re/im planes [B, 2, F, T] -> audio [B, hop*(T-1)] — base.py:39-47 (``InverseSpectrogram``, length=None)."""
    spec = torch.view_as_complex(planes.permute(0, 2, 3, 1).contiguous())
    return torch.istft(spec, cfg.n_fft, hop_length=cfg.hop, win_length=cfg.win_length, window=window,
                       center=True, normalized=False, onesided=True, length=None, return_complex=False)


# --------------------------------------------------------------------------- patch (de)embedding and resampling
def patch_embed(sd: Dict[str, Tensor], planes: Tensor, cfg: OracleConfig) -> Tuple[Tensor, Tuple[int, int]]:
    """[B,2,F,T] -> tokens [B, H*W, C0] (freq-major) — scale.py:42-50, base.py:149-150."""
    pf, pt = cfg.patch_size
    y = F.conv2d(planes, sd["encoder.patch_embed.proj.weight"], sd["encoder.patch_embed.proj.bias"], stride=(pf, pt))
    H, W = y.shape[2], y.shape[3]
    tok = y.flatten(2).transpose(1, 2)
    tok = F.layer_norm(tok, (tok.shape[-1],), sd["encoder.patch_embed.norm.weight"], sd["encoder.patch_embed.norm.bias"])
    return tok, (H, W)


def patch_merge(sd: Dict[str, Tensor], prefix: str, x: Tensor, H: int) -> Tensor:
    """Pair frequency rows (2h, 2h+1) channel-wise, LN(2C), bias-free Linear — scale.py:7-14,97-115."""
    B, L, C = x.shape
    W = L // H
    g = x.view(B, H, W, C)
    if H % 2:
        g = F.pad(g, (0, 0, 0, 0, 0, 1))
        H += 1
    g = g.view(B, H // 2, 2, W, C).permute(0, 1, 3, 2, 4).reshape(B, (H // 2) * W, 2 * C)
    g = F.layer_norm(g, (2 * C,), sd[f"{prefix}.norm.weight"], sd[f"{prefix}.norm.bias"])
    return F.linear(g, sd[f"{prefix}.down.weight"])


def patch_split(sd: Dict[str, Tensor], prefix: str, x: Tensor, H: int) -> Tensor:
    """LN(C), bias-free Linear(C -> 2C'), first C' to row 2h, last C' to row 2h+1 — scale.py:16-23,131-145."""
    B, L, C = x.shape
    W = L // H
    y = F.linear(F.layer_norm(x, (C,), sd[f"{prefix}.norm.weight"], sd[f"{prefix}.norm.bias"]), sd[f"{prefix}.up.weight"])
    Co = y.shape[-1] // 2
    y = y.view(B, H, W, 2, Co).permute(0, 1, 3, 2, 4).reshape(B, 2 * H * W, Co)
    return y


def patch_deembed(sd: Dict[str, Tensor], tok: Tensor, cfg: OracleConfig) -> Tensor:
    """tokens [B, H*W, C0] -> planes [B, 2, F, T] — scale.py:73-81 (conv5x5, pixel-shuffle (3,2), conv3x3)."""
    pf, pt = cfg.patch_size
    B, L, C = tok.shape
    H = cfg.top_freq
    W = L // H
    y = tok.view(B, H, W, C).permute(0, 3, 1, 2)
    y = F.conv2d(y, sd["decoder.patch_deembed.de_proj1.weight"], sd["decoder.patch_deembed.de_proj1.bias"], padding=2)
    # channel n = (s1*pt + s2)*C + c  ->  pixel (pf*h + s1, pt*w + s2), channel c
    y = y.permute(0, 2, 3, 1).reshape(B, H, W, pf, pt, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H * pf, W * pt, C)
    y = F.conv2d(y.permute(0, 3, 1, 2), sd["decoder.patch_deembed.de_proj2.weight"],
                 sd["decoder.patch_deembed.de_proj2.bias"], padding=1)
    return y


# --------------------------------------------------------------------------- Swin blocks
def _regions(P: int, ws: int, shift: int) -> Tensor:
    """Region id along one axis of the padded, *shifted* map: attention.py:59-64."""
    r = torch.zeros(P, dtype=torch.long)
    r[P - ws:P - shift] = 1
    r[P - shift:] = 2
    return r


def shifted_window_mask(Hp: int, Wp: int, ws: int) -> Tensor:
    """[nW, ws*ws, ws*ws] additive mask, 0 where two tokens share a region else -100 — attention.py:56-75."""
    shift = ws // 2
    ids = (_regions(Hp, ws, shift)[:, None] * 3 + _regions(Wp, ws, shift)[None, :]).to(torch.float32)
    win = ids.view(Hp // ws, ws, Wp // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = win[:, None, :] - win[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def _to_windows(g: Tensor, ws: int) -> Tensor:
    B, Hp, Wp, C = g.shape
    return g.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)


def _from_windows(w: Tensor, ws: int, B: int, Hp: int, Wp: int) -> Tensor:
    C = w.shape[-1]
    return w.view(B, Hp // ws, Wp // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)


def window_attention(sd: Dict[str, Tensor], prefix: str, xw: Tensor, heads: int, mask: Optional[Tensor]) -> Tensor:
    """Multi-head attention inside each 4x4 window — attention.py:215-244."""
    Bw, N, C = xw.shape
    hd = C // heads
    qkv = F.linear(xw, sd[f"{prefix}.qkv.weight"], sd[f"{prefix}.qkv.bias"])
    qkv = qkv.reshape(Bw, N, 3, heads, hd).permute(2, 0, 3, 1, 4).contiguous()
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q * (hd ** -0.5)) @ k.transpose(-2, -1)
    table = sd[f"{prefix}.relative_position_bias_table"]
    index = sd[f"{prefix}.relative_position_index"]
    bias = table[index.reshape(-1)].view(N, N, heads).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(Bw // nW, nW, heads, N, N) + mask[None, :, None]).view(-1, heads, N, N)
    attn = torch.softmax(attn, dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(Bw, N, C)
    return F.linear(out, sd[f"{prefix}.proj.weight"], sd[f"{prefix}.proj.bias"])


def swin_block(sd: Dict[str, Tensor], prefix: str, x: Tensor, H: int, W: int, heads: int, ws: int,
               shift: int, mask: Optional[Tensor]) -> Tensor:
    """One (shifted-)window block — attention.py:129-178.  Zero padding happens AFTER norm1."""
    B, L, C = x.shape
    y = F.layer_norm(x, (C,), sd[f"{prefix}.norm1.weight"], sd[f"{prefix}.norm1.bias"]).view(B, H, W, C)
    pad_w, pad_h = (-W) % ws, (-H) % ws
    y = F.pad(y, (0, 0, 0, pad_w, 0, pad_h))
    Hp, Wp = H + pad_h, W + pad_w
    if shift:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
    a = window_attention(sd, f"{prefix}.attn", _to_windows(y, ws), heads, mask if shift else None)
    y = _from_windows(a, ws, B, Hp, Wp)
    if shift:
        y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
    y = y[:, :H, :W, :].contiguous().view(B, L, C)
    x = x + y
    h = F.layer_norm(x, (C,), sd[f"{prefix}.norm2.weight"], sd[f"{prefix}.norm2.bias"])
    h = F.gelu(F.linear(h, sd[f"{prefix}.mlp.linear_1.weight"], sd[f"{prefix}.mlp.linear_1.bias"]))
    h = F.linear(h, sd[f"{prefix}.mlp.linear_2.weight"], sd[f"{prefix}.mlp.linear_2.bias"])
    return x + h


def swin_layer(sd: Dict[str, Tensor], prefix: str, x: Tensor, H: int, W: int, heads: int, depth: int, ws: int,
               scale: Optional[str]) -> Tuple[Tensor, int, int]:
    """``depth`` alternating W-MSA / SW-MSA blocks then the optional resample — attention.py:48-91."""
    Hp, Wp = math.ceil(H / ws) * ws, math.ceil(W / ws) * ws
    mask = shifted_window_mask(Hp, Wp, ws)
    for j in range(depth):
        x = swin_block(sd, f"{prefix}.swint_blocks.{j}", x, H, W, heads, ws, (ws // 2) if j % 2 else 0, mask)
    if scale == "down":
        return patch_merge(sd, f"{prefix}.subsample", x, H), (H + 1) // 2, W
    if scale == "up":
        return patch_split(sd, f"{prefix}.subsample", x, H), H * 2, W
    return x, H, W


# --------------------------------------------------------------------------- product VQ
def pvq_frames(z: Tensor, in_freq: int, overlap: int) -> Tensor:
    """tokens [B, H*W, C] -> VQ frames [B, W/overlap, overlap*C*H], inner order (o, c, h) — quantization.py:388-409."""
    B, L, C = z.shape
    W = L // in_freq
    if W % overlap:
        raise AssertionError("Time dimension must be multiple of overlap")
    f = z.view(B, in_freq, W, C).permute(0, 2, 3, 1).reshape(B, W, C * in_freq)
    return f.reshape(B, W // overlap, overlap * C * in_freq)


def pvq_unframes(f: Tensor, in_freq: int, overlap: int) -> Tensor:
    """inverse of :func:`pvq_frames` — quantization.py:411-432."""
    B, T, D = f.shape
    C = D // (overlap * in_freq)
    return f.reshape(B, T * overlap, C, in_freq).permute(0, 3, 1, 2).reshape(B, in_freq * T * overlap, C)


def codebook_argmin(z: Tensor, table: Tensor, l2norm: bool = True) -> Tensor:
    """THE RVQ argmin.  z [B, T, d], table [K, d] -> indices [B, T] int64 — codebook.py:20-43.

    Both sides are L2-normalised (eps 1e-12); the distance is ``|z|^2 - (2 z).c + |c|^2`` in that
    association; ``min`` returns the first minimum.
    """
    B, T, d = z.shape
    zf = z.reshape(B * T, d)
    cb = table
    if l2norm:
        cb = F.normalize(cb, dim=-1)
        zf = F.normalize(zf, dim=-1)
    dist = zf.pow(2).sum(1, keepdim=True) - 2 * zf @ cb.t() + cb.pow(2).sum(1, keepdim=True).t()
    return dist.min(1).indices.view(B, T)


def pvq_project(sd: Dict[str, Tensor], prefix: str, z: Tensor, in_freq: int, cfg: OracleConfig) -> List[Tensor]:
    """Per-group bias-free down-projection of the VQ frames — quantization.py:80-86,120."""
    f = pvq_frames(z, in_freq, cfg.overlap)
    dims = _split_dimension(f.shape[-1], cfg.group_size)
    out, s = [], 0
    for g, n in enumerate(dims):
        out.append(F.linear(f[..., s:s + n], sd[f"{prefix}.down_projs.{g}.weight"]))
        s += n
    return out


def pvq_encode(sd: Dict[str, Tensor], prefix: str, z: Tensor, in_freq: int, cfg: OracleConfig) -> Tensor:
    """tokens -> codes [B, groups, T] — quantization.py:74-91."""
    ze = pvq_project(sd, prefix, z, in_freq, cfg)
    return torch.stack([codebook_argmin(ze[g], sd[f"{prefix}.vqs.{g}.embedding.weight"], cfg.l2norm)
                        for g in range(cfg.group_size)], dim=1)


def pvq_decode(sd: Dict[str, Tensor], prefix: str, codes: Tensor, in_freq: int, cfg: OracleConfig) -> Tensor:
    """codes [B, groups, T] -> tokens; gathers the UN-normalised table — quantization.py:93-108,124-136; codebook.py:45-55."""
    parts = []
    for g in range(cfg.group_size):
        zq = F.embedding(codes[:, g, :], sd[f"{prefix}.vqs.{g}.embedding.weight"])
        parts.append(F.linear(zq, sd[f"{prefix}.up_projs.{g}.weight"]))
    return pvq_unframes(torch.cat(parts, dim=-1), in_freq, cfg.overlap)


def pvq_forward_eval(sd: Dict[str, Tensor], prefix: str, z: Tensor, in_freq: int, cfg: OracleConfig):
    """Eval-mode ``ProductVectorQuantize.forward`` incl. the MSE "losses" — quantization.py:31-72, codebook.py:57-75."""
    ze = pvq_project(sd, prefix, z, in_freq, cfg)
    codes, parts, loss = [], [], 0.0
    for g in range(cfg.group_size):
        table = sd[f"{prefix}.vqs.{g}.embedding.weight"]
        code = codebook_argmin(ze[g], table, cfg.l2norm)
        zq = F.embedding(code, table)
        loss = loss + F.mse_loss(zq, ze[g], reduction="none").mean([1, 2])
        parts.append(F.linear(zq, sd[f"{prefix}.up_projs.{g}.weight"]))
        codes.append(code)
    z_q = pvq_unframes(torch.cat(parts, dim=-1), in_freq, cfg.overlap)
    loss = loss / cfg.group_size
    return z_q, torch.stack(codes, dim=1), loss, loss


# --------------------------------------------------------------------------- the codec
class EscOracle:
    """encode / decode / forward(eval) of the reference ``ESC`` on a plain state dict."""

    def __init__(self, cfg_kwargs: dict, state_dict: Dict[str, Tensor]):
        self.cfg = OracleConfig(**cfg_kwargs)
        self.sd = {k: v.detach().to("cpu") for k, v in state_dict.items()}
        self.taps: Dict[str, Tensor] = {}
        self.record_taps = False

    # -- helpers
    def _tap(self, name: str, t: Tensor) -> None:
        if self.record_taps:
            self.taps[name] = t.detach().clone()

    def _layer(self, prefix: str, x: Tensor, H: int, W: int, heads: int, scale: Optional[str]):
        c = self.cfg
        return swin_layer(self.sd, prefix, x, H, W, heads, c.swin_depth, c.window_size, scale)

    # -- encoder: base.py:143-158
    def encoder(self, planes: Tensor) -> Tuple[List[Tensor], Tuple[int, int]]:
        c = self.cfg
        x, (H, W) = patch_embed(self.sd, planes, c)
        self._tap("patch_embed", x)
        x, H, W = self._layer("encoder.pre_nn", x, H, W, c.swin_heads[0], None)
        enc_hs = [x]
        for i in range(len(c.h_dims) - 1):
            x, H, W = self._layer(f"encoder.blocks.{i}", x, H, W, c.swin_heads[i], "down")
            enc_hs.append(x)
        for i, t in enumerate(enc_hs):
            self._tap(f"enc_hs.{i}", t)
        return enc_hs, (H, W)

    def spec_transform(self, x: Tensor) -> Tensor:
        return stft_planes(x, self.sd["ft.window"], self.cfg)

    def audio_reconstruct(self, planes: Tensor) -> Tensor:
        return istft_audio(planes, self.sd["ift.window"], self.cfg)

    # -- csrvq.py:131-158
    @torch.no_grad()
    def encode(self, x: Tensor, num_streams: int = 6) -> Tuple[Tensor, Tuple[int, int]]:
        c, sd = self.cfg, self.sd
        planes = self.spec_transform(x)
        self._tap("stft", planes)
        enc_hs, (H, W) = self.encoder(planes)
        feat_shape = (H, W)
        q0 = c.quantizer_geometry(0)[1]
        code0 = pvq_encode(sd, "quantizers.0", enc_hs[-1], q0, c)
        if num_streams == 1:
            return code0.unsqueeze(1), feat_shape
        dec = pvq_decode(sd, "quantizers.0", code0, q0, c)
        codes = [code0]
        for i in range(num_streams - 1):
            qf = c.quantizer_geometry(i + 1)[1]
            residual = enc_hs[-1 - i] - dec
            self._tap(f"residual.{i + 1}", residual)
            code = pvq_encode(sd, f"quantizers.{i + 1}", residual, qf, c)
            codes.append(code)
            if len(codes) == num_streams:
                break
            refine = pvq_decode(sd, f"quantizers.{i + 1}", code, qf, c) + dec
            dec, H, W = self._layer(f"decoder.blocks.{i}", refine, H, W, c.dec_heads[i], "up")
        return torch.stack(codes, dim=1), feat_shape

    # -- csrvq.py:160-182 + codecs.py:83-94
    @torch.no_grad()
    def decode_features(self, codes: Tensor, feat_shape: Tuple[int, int]) -> Tensor:
        c, sd = self.cfg, self.sd
        S = codes.shape[1]
        H, W = feat_shape
        dec = pvq_decode(sd, "quantizers.0", codes[:, 0], c.quantizer_geometry(0)[1], c)
        self._tap("dec_hs.0", dec)
        for i in range(len(c.h_dims) - 1):
            if i < S - 1:
                dec = pvq_decode(sd, f"quantizers.{i + 1}", codes[:, i + 1], c.quantizer_geometry(i + 1)[1], c) + dec
            dec, H, W = self._layer(f"decoder.blocks.{i}", dec, H, W, c.dec_heads[i], "up")
            self._tap(f"dec_hs.{i + 1}", dec)
        dec, H, W = self._layer("decoder.post_nn", dec, H, W, c.dec_heads[-1], None)
        self._tap("post_nn", dec)
        planes = patch_deembed(sd, dec, c)
        self._tap("recon_feat", planes)
        return planes

    @torch.no_grad()
    def decode(self, codes: Tensor, feat_shape: Tuple[int, int] = (2, 1000)) -> Tensor:
        return self.audio_reconstruct(self.decode_features(codes, feat_shape))

    # -- eval-mode forward: codecs.py:30-66, csrvq.py:23-48,97-129
    @torch.no_grad()
    def forward(self, x: Tensor, x_feat: Optional[Tensor] = None, num_streams: int = 6) -> dict:
        c, sd = self.cfg, self.sd
        planes = self.spec_transform(x) if x_feat is None else x_feat.permute(0, 3, 1, 2)
        enc_hs, (H, W) = self.encoder(planes)
        dec, code, cm, cb = pvq_forward_eval(sd, "quantizers.0", enc_hs[-1] - 0.0, c.quantizer_geometry(0)[1], c)
        dec = dec + 0.0
        codes = [code]
        cm_loss, cb_loss = cm, cb
        for i in range(len(c.h_dims) - 1):
            if i < num_streams - 1:
                zq, code, cm, cb = pvq_forward_eval(sd, f"quantizers.{i + 1}", enc_hs[-1 - i] - dec,
                                                    c.quantizer_geometry(i + 1)[1], c)
                dec = zq + dec
                codes.append(code)
                cm_loss = cm_loss + cm
                cb_loss = cb_loss + cb
            dec, H, W = self._layer(f"decoder.blocks.{i}", dec, H, W, c.dec_heads[i], "up")
        dec, H, W = self._layer("decoder.post_nn", dec, H, W, c.dec_heads[-1], None)
        recon_feat = patch_deembed(sd, dec, c)
        return {"cm_loss": cm_loss, "cb_loss": cb_loss, "raw_audio": x,
                "recon_audio": self.audio_reconstruct(recon_feat), "raw_feat": planes,
                "recon_feat": recon_feat, "codes": torch.stack(codes, dim=1)}


# --------------------------------------------------------------------------- RVQCodecs (the reference's RVQ ablation codec)
def rvq_project(sd: Dict[str, Tensor], z: Tensor, in_freq: int, cfg: OracleConfig) -> List[Tensor]:
    """pre_process + per-group ``proj_down`` of ProductResidualVectorQuantize — quantization.py:343-353, 388-409."""
    f = pvq_frames(z, in_freq, cfg.overlap)
    dims = _split_dimension(f.shape[-1], cfg.group_size)
    out, s = [], 0
    for m, n in enumerate(dims):
        out.append(F.linear(f[..., s:s + n], sd[f"quantizers.vqs.{m}.proj_down.weight"]))
        s += n
    return out


def rvq_chain(sd: Dict[str, Tensor], m: int, z: Tensor, num_streams: int, cfg: OracleConfig):
    """One group's residual chain in eval mode — ResidualVectorQuantize.residual_vector_quantization, quantization.py:170-195
    (== quantize_to_code :223-237 for the codes).  Returns (z_q, codes [B, S, T], loss [B])."""
    z_q, codes, loss = 0.0, [], 0.0
    residual = z
    for i in range(num_streams):
        table = sd[f"quantizers.vqs.{m}.vqs.{i}.embedding.weight"]
        code = codebook_argmin(residual, table, cfg.l2norm)
        z_q_i = F.embedding(code, table)
        loss = loss + F.mse_loss(z_q_i, residual, reduction="none").mean([1, 2])      # codebook.py:71-73 (eval)
        residual = residual - z_q_i
        z_q = z_q + z_q_i
        codes.append(code)
    return z_q, torch.stack(codes, dim=1), loss


class RvqOracle(EscOracle):
    """encode / decode / forward(eval) of the reference ``RVQCodecs`` (codecs.py:96-181) on a plain state dict."""

    def __init__(self, cfg_kwargs: dict, state_dict: Dict[str, Tensor]):
        kw = dict(cfg_kwargs)
        self.num_rvqs = int(kw.pop("num_rvqs", 6))
        d = kw.pop("codebook_dim", 8)
        super().__init__(dict(kw, codebook_dims=[d] * kw.get("max_streams", 6)), state_dict)

    def _quantize(self, z: Tensor, num_streams: int):
        c = self.cfg
        H0 = c.quantizer_geometry(0)[1]
        ze = rvq_project(self.sd, z, H0, c)
        parts, codes, loss = [], [], 0.0
        for m in range(c.group_size):
            zq_m, codes_m, loss_m = rvq_chain(self.sd, m, ze[m], num_streams, c)
            parts.append(zq_m)
            codes.append(codes_m)
            loss = loss + loss_m
        return parts, torch.stack(codes, dim=2), loss / c.group_size            # codes [B, S, groups, T]

    def _dequantize(self, parts: List[Tensor]) -> Tensor:
        c = self.cfg
        ups = [F.linear(p, self.sd[f"quantizers.vqs.{m}.proj_up.weight"]) for m, p in enumerate(parts)]
        return pvq_unframes(torch.cat(ups, dim=-1), c.quantizer_geometry(0)[1], c.overlap)

    def _decoder(self, z_q: Tensor, feat_shape: Tuple[int, int]) -> Tensor:
        """Decoder.forward — base.py:194-203."""
        c = self.cfg
        H, W = feat_shape
        for i in range(len(c.h_dims) - 1):
            z_q, H, W = self._layer(f"decoder.blocks.{i}", z_q, H, W, c.dec_heads[i], "up")
        z_q, H, W = self._layer("decoder.post_nn", z_q, H, W, c.dec_heads[-1], None)
        return patch_deembed(self.sd, z_q, c)

    @torch.no_grad()
    def encode(self, x: Tensor, num_streams: int = 6):
        enc_hs, feat_shape = self.encoder(self.spec_transform(x))
        return self._quantize(enc_hs[-1], num_streams)[1], feat_shape

    @torch.no_grad()
    def decode(self, codes: Tensor, feat_shape: Tuple[int, int]) -> Tensor:
        c = self.cfg
        parts = []
        for m in range(c.group_size):                                          # dequantize_code, quantization.py:239-245
            z_q = 0.0
            for i in range(codes.shape[1]):
                z_q = z_q + F.embedding(codes[:, i, m, :], self.sd[f"quantizers.vqs.{m}.vqs.{i}.embedding.weight"])
            parts.append(z_q)
        return self.audio_reconstruct(self._decoder(self._dequantize(parts), feat_shape))

    @torch.no_grad()
    def forward(self, x: Tensor, x_feat: Optional[Tensor] = None, num_streams: int = 6) -> dict:
        planes = self.spec_transform(x) if x_feat is None else x_feat.permute(0, 3, 1, 2)
        enc_hs, feat_shape = self.encoder(planes)
        parts, codes, loss = self._quantize(enc_hs[-1], num_streams)
        recon_feat = self._decoder(self._dequantize(parts), feat_shape)
        return {"cm_loss": loss, "cb_loss": loss, "raw_audio": x, "recon_audio": self.audio_reconstruct(recon_feat),
                "raw_feat": planes, "recon_feat": recon_feat, "codes": codes}
