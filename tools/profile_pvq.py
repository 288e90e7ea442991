"""BASELINE configs[3] under a profiler: the six fused RVQ stream steps (escb_pvq_stream) on 1024 VQ frames.
usage: python tools/profile_pvq.py [iters=3] [W=2048] [B=1]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-speech-codec_b200"))
sys.path.insert(0, ROOT)
import torch
from bench import BASE
from escb200 import native
from escb200.codec import ESC
from escb200.spec import CodecSpec
from escb200.synthetic import synth_state_dict

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
m = ESC(**BASE)
m.load_state_dict(synth_state_dict(CodecSpec.from_kwargs(**BASE), 0))
m = m.eval().to(dev)
lib, h, spec = native.lib(), m._handle(dev), m.spec
ws = m._ws(dev, h.workspace_bytes(B, W))
st = m._stream(dev)
g = torch.Generator().manual_seed(100)
streams = []
for q, qs in enumerate(spec.quantizers()):
    enc = torch.randn(B, qs.in_freq * W, qs.in_dim, generator=g).to(dev)
    dec = None if q == 0 else torch.randn(B, qs.in_freq * W, qs.in_dim, generator=g).to(dev)
    streams.append((q, enc, dec, torch.empty_like(enc), torch.empty((B, 3, W // 2), dtype=torch.int64, device=dev)))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for it in range(iters):
    flush.zero_()
    for q, enc, dec, out, codes in streams:
        native.check(lib.escb_pvq_stream(h.ptr, q, native.ptr(enc), native.ptr(dec), B, W, native.ptr(codes), native.ptr(out),
                                         native.ptr(ws), ws.numel(), st))
torch.cuda.synchronize()
print("ok")
