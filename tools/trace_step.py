"""Per-launch role breakdown of the tcgen05 GEMM engine (debug): needs tools/libescb200_trace.so, a build of
csrc/ with -DESCB_TC_TRACE.  usage: python tools/trace_step.py [batch]"""
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

B = int(sys.argv[1]) if len(sys.argv) > 1 else 36
m = ESC(**BASE)
m.load_state_dict(synth_state_dict(CodecSpec.from_kwargs(**BASE), 0))
m = m.eval().cuda()
x = synth_audio(B, 48000, seed=1000).cuda()
codes, fs = m.encode(x, 6)
audio = m.decode(codes, fs)
torch.cuda.synchronize()
h = m._handle(torch.device("cuda", 0))
L = native.lib()
L.escb_debug_trace.argtypes = [C.c_void_p, C.c_void_p]
L.escb_debug_trace.restype = C.c_int
hp = h._h if hasattr(h, "_h") else h.ptr
L.escb_debug_trace(hp, None)          # allocate
for phase in ("encode", "decode"):
    if phase == "encode":
        codes, fs = m.encode(x, 6)
    else:
        audio = m.decode(codes, fs)
    torch.cuda.synchronize()
    buf = np.zeros(1024 * 16, dtype=np.uint64)
    L.escb_debug_trace(hp, buf.ctypes.data_as(C.c_void_p))
    t = buf.reshape(1024, 16).astype(np.float64)
    print(f"== {phase}: per launch: N K BN nsub res | tiles ctas | kclk/cta | epi wait% | prod: wait_a_empty% convert% issue% | "
          f"mma: wait_a_full% wait_acc_empty% wait_b_full% issue% | prod tile-init%")
    for i in range(1024):
        n = t[i, 13]
        sig = int(buf.reshape(1024, 16)[i, 15])
        if sig >> 60 == 0xF:        # fused MLP kernel (mlp_fused.cu)
            C_, ntl, ctas = sig & 0xFFFFF, (sig >> 20) & 0xFFFFFFFF, max(t[i, 14], 1)
            pct = lambda k, tot: 100 * t[i, k] / max(t[i, tot], 1)
            print(f"{i:3d} MLP-FUSED C={C_:3d} | {ntl:5d} {int(ctas):3d} | {t[i,0]/ctas/1e3:7.1f} | mma wait: a1_full {pct(1,0):3.0f} h_full {pct(2,0):3.0f} "
                  f"acc_free {pct(3,0):3.0f} w_full {pct(4,0):3.0f} | ln wait: x_full {pct(6,5):3.0f} a1_free {pct(7,5):3.0f} | "
                  f"gelu wait: r_full {pct(9,8):3.0f} l_free {pct(10,8):3.0f} | out wait: acc_full {pct(12,11):3.0f}")
            continue
        if n == 0:
            continue
        sig = int(buf.reshape(1024, 16)[i, 15])
        N, K, BN, nsub, res = sig >> 40, (sig >> 20) & 0xFFFFF, (sig >> 8) & 0xFFF, (sig >> 4) & 0xF, sig & 0xF
        tot = t[i, 9] / n
        pe = 100 * t[i, 1] / max(t[i, 9], 1)
        pw, pc, pi = (100 * t[i, k] / max(t[i, 10], 1) for k in (2, 3, 4))
        ma, mc, mb, mi = (100 * t[i, k] / max(t[i, 11], 1) for k in (5, 6, 7, 0))
        lw = 100 * t[i, 12] / max(t[i, 10], 1)   # producer: per-tile init + prefetch share
        print(f"{i:3d} N={N:4d} K={K:4d} BN={BN:3d}x{nsub} r{res} | {int(t[i,14]):5d} {int(n):3d} | {tot/1e3:7.1f} | {pe:4.0f} | "
              f"{pw:4.0f} {pc - pw:4.0f} {pi:4.0f} | {ma:4.0f} {mc:4.0f} {mb:4.0f} {mi:4.0f} | {lw:4.0f}")
