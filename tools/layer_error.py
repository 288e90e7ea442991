"""Per-layer numerical error of the CUDA engines against a float64 evaluation of the oracle: which TransformerLayer
loses precision on the tcgen05 (3xTF32) path?  usage: python tools/layer_error.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-speech-codec_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import BASE, Unit, make_native
from oracle.esc_oracle import OracleConfig, swin_layer
from escb200.spec import CodecSpec
from escb200.synthetic import synth_state_dict

sd = synth_state_dict(CodecSpec.from_kwargs(**BASE), 0)
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
c = OracleConfig(**BASE)
L, W, B = 6, 300, 2
g = torch.Generator().manual_seed(5)
layers = []
for li in range(2 * L):
    if li == 0:
        layers.append(("encoder.pre_nn", c.h_dims[0], c.swin_heads[0], None, 64))
    elif li < L:
        i = li - 1
        layers.append((f"encoder.blocks.{i}", c.h_dims[i], c.swin_heads[i], "down", 64 >> i))
    elif li < 2 * L - 1:
        i = li - L
        layers.append((f"decoder.blocks.{i}", c.dec_h_dims[i], c.dec_heads[i], "up", 2 << i))
    else:
        layers.append(("decoder.post_nn", c.h_dims[0], c.dec_heads[-1], None, 64))
xs = [torch.randn(B, H * W, C, generator=g) for (_, C, _, _, H) in layers]
refs = []
for (prefix, C, heads, scale, H), x in zip(layers, xs):
    r, _, _ = swin_layer(sd64, prefix, x.double(), H, W, heads, c.swin_depth, c.window_size, scale)
    refs.append(r)
for env in ({"ESCB_GEMM": "simt"}, {}, {"ESCB_FUSE_MLP": "0"}):
    for k in list(os.environ):
        if k.startswith("ESCB_"):
            del os.environ[k]
    os.environ.update(env)
    m, _ = make_native(BASE, 0)
    u = Unit(m)
    errs = []
    for li, ((prefix, C, heads, scale, H), x, r) in enumerate(zip(layers, xs, refs)):
        y = u.swin_layer(li, x, H, W, tuple(r.shape)).double()
        errs.append(float((y - r).abs().max() / r.abs().max()))
    print(f"{str(env):28s} max-abs error / max|ref| per layer: " + " ".join(f"{e:.1e}" for e in errs))
