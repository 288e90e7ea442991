"""Accuracy / speed of the tcgen05 engine per accumulator-split policy (ESCB_ACC="nmain,corr", read at pack time):
per-layer error against float64, code flips against the REAL reference (baseline/_ref, CPU fp32) and ms per
encode+decode step.  usage: python tools/acc_sweep.py [batches=2] [env ...]   e.g.  ESCB_ACC=1,1  ESCB_ACC=2,1:ESCB_FUSE_MLP=0
(GPU box; needs the staged reference)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-speech-codec_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bench import BASE
from helpers import Unit, make_native
from escb200.codec import ESC
from escb200.spec import CodecSpec
from escb200.synthetic import synth_audio, synth_state_dict
from oracle import ref_loader
from oracle.esc_oracle import OracleConfig, swin_layer

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2
envs = [dict(kv.split("=") for kv in a.split(":") if kv) for a in sys.argv[2:]] or [{}]
sd = synth_state_dict(CodecSpec.from_kwargs(**BASE), 0)
torch.set_num_threads(os.cpu_count() or 1)
ref = ref_loader.make_reference_model(BASE, sd)
xs, rcs = [], []
with torch.no_grad():
    for b in range(nb):
        x = synth_audio(36, 48000, seed=1000 + b)
        rc, _ = ref.encode(x, 6)
        xs.append(x)
        rcs.append(rc)

# per-layer inputs / float64 references (tools/layer_error.py)
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
lx = [torch.randn(B, H * W, C, generator=g) for (_, C, _, _, H) in layers]
lref = [swin_layer(sd64, p, x.double(), H, W, h, c.swin_depth, c.window_size, s)[0] for (p, C, h, s, H), x in zip(layers, lx)]

for env in envs:
    for k in list(os.environ):
        if k.startswith("ESCB_"):
            del os.environ[k]
    os.environ.update(env)
    m, _ = make_native(BASE, 0)
    u = Unit(m)
    errs = []
    for li, ((p, C, h, s, H), x, r) in enumerate(zip(layers, lx, lref)):
        y = u.swin_layer(li, x, H, W, tuple(r.shape)).double()
        errs.append(float((y - r).abs().max() / r.abs().max()))
    model = ESC(**BASE)
    model.load_state_dict(sd)
    model = model.eval().cuda()
    flips, total_bad = [], 0
    for b in range(nb):
        cc, f = model.encode(xs[b].cuda(), 6)
        bad = cc.cpu() != rcs[b]
        total_bad += int(bad.sum())
        for cl in torch.nonzero(bad.flatten(1).sum(1)).flatten().tolist():
            flips.append((36 * b + cl, int(torch.nonzero(bad[cl].flatten(1).sum(1)).flatten()[0])))
    xg = xs[0].cuda()
    for _ in range(3):
        cc, f = model.encode(xg, 6)
        model.decode(cc, f)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        cc, f = model.encode(xg, 6)
        model.decode(cc, f)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{str(env):44s} {ms:6.2f} ms/step | clips with a flip {len(flips):2d} of {36 * nb} (codes differing {total_bad}) {flips} | "
          f"layer err " + " ".join(f"{e * 1e6:.1f}" for e in errs) + " e-6", flush=True)
