"""CTA-0 event timeline of the fused MLP kernel (trace build, tools/libescb200_trace.so):
usage: python tools/timeline_mlp.py [C=45] [batch=36] -> prints events sorted by clock (first launch with that C)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-speech-codec_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from escb200 import native

native.library_path = lambda: os.path.join(ROOT, "tools", "libescb200_trace.so")
from bench import BASE
from escb200.codec import ESC
from escb200.spec import CodecSpec
from escb200.synthetic import synth_audio, synth_state_dict

Cc = int(sys.argv[1]) if len(sys.argv) > 1 else 45
B = int(sys.argv[2]) if len(sys.argv) > 2 else 36
m = ESC(**BASE)
m.load_state_dict(synth_state_dict(CodecSpec.from_kwargs(**BASE), 0))
m = m.eval().cuda()
x = synth_audio(B, 48000, seed=1000).cuda()
m.encode(x, 1)
torch.cuda.synchronize()
L = native.lib()
L.escb_debug_timeline.argtypes = [C.c_int, C.c_void_p, C.c_int]
L.escb_debug_timeline.restype = C.c_int
L.escb_debug_timeline(Cc, None, 0)
m.encode(x, 1)
torch.cuda.synchronize()
buf = np.zeros(8192, dtype=np.uint64)
n = L.escb_debug_timeline(0, buf.ctypes.data_as(C.c_void_p), 8192)
ev = [(int(v) & 0xFFFFFFFFFF, int(v) >> 56, (int(v) >> 40) & 0xFFFF) for v in buf[:n]]
ev.sort()
names = {1: "G1 begin", 2: "G1 issued", 3: "G2 begin", 4: "G2 issued", 5: "GELU begin", 6: "GELU end", 7: "LN begin",
         8: "LN end", 9: "OUT begin", 10: "OUT end", 11: "  mma stage ready", 12: "  mma lane0 past issue", 13: "  mma warp synced"}
t0 = ev[0][0] if ev else 0
print(f"{n} events, C={Cc}")
for clk, e, a in ev[:int(sys.argv[3]) if len(sys.argv) > 3 else 400]:
    print(f"{clk - t0:9d}  {names.get(e, e):22s} {a}")
