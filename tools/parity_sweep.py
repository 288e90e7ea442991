"""Code-index parity against the REAL reference (baseline/_ref, CPU fp32) over many clips, per engine variant.
usage: python tools/parity_sweep.py [clips=36] [seed=1000]   (GPU box; needs the staged reference)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-speech-codec_b200"))
sys.path.insert(0, ROOT)
import torch
from bench import BASE
from escb200.codec import ESC
from escb200.spec import CodecSpec
from escb200.synthetic import synth_audio, synth_state_dict
from oracle import ref_loader

n = int(sys.argv[1]) if len(sys.argv) > 1 else 36
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
sd = synth_state_dict(CodecSpec.from_kwargs(**BASE), 0)
x = synth_audio(n, 48000, seed=seed)
torch.set_num_threads(os.cpu_count() or 1)
ref = ref_loader.make_reference_model(BASE, sd)
with torch.no_grad():
    rc, fs = ref.encode(x, 6)
    ra = ref.decode(rc, fs)
    ref64 = ref.double()
    rc64, _ = ref64.encode(x.double(), 6)
print(f"reference fp32 vs fp64 (CPU): {(rc != rc64).sum().item()} code mismatches of {rc.numel()}")


def run(env):
    for k in list(os.environ):
        if k.startswith("ESCB_"):
            del os.environ[k]
    os.environ.update(env)
    m = ESC(**BASE)
    m.load_state_dict(sd)
    m = m.eval().cuda()
    c, f = m.encode(x.cuda(), 6)
    a = m.decode(c, f)
    c = c.cpu()
    bad = (c != rc)
    per_clip = bad.flatten(1).sum(1)
    first = []
    for b in torch.nonzero(per_clip).flatten().tolist():
        s = int(torch.nonzero(bad[b].flatten(1).sum(1)).flatten()[0])
        first.append((b, s, int(bad[b, s].sum())))
    print(f"{str(env):60s} mismatches vs ref fp32 {int(bad.sum()):5d} | vs ref fp64 {int((c != rc64).sum()):5d} | "
          f"audio max-abs {float((a.cpu() - ra).abs().max()):.2e} | (clip, first stream, n) {first}")


for env in ({}, {"ESCB_GEMM": "simt"}, {"ESCB_FUSE_MLP": "0"}, {"ESCB_FUSE_PVQ": "0"}, {"ESCB_FUSE_ATTN_MAXC": "0"},
            {"ESCB_FUSE_MLP": "0", "ESCB_FUSE_PVQ": "0", "ESCB_EMIT_STATS": "0"}):
    run(env)
