"""Per-launch CUDA-event times of one encode+decode step (ESCB_PROFILE_DUMP), grouped by (kernel class, algorithmic
flops) = one row per class and level.  usage: python tools/per_launch.py [batch=36] [class substring]"""
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("_PL_CHILD"):
    sys.path.insert(0, os.path.join(ROOT, "efficient-speech-codec_b200"))
    sys.path.insert(0, ROOT)
    import torch
    if os.environ.get("PL_LIB"):                       # another build of the library (csrc/Makefile `variant`)
        from escb200 import native
        native.library_path = lambda: os.path.abspath(os.environ["PL_LIB"])
    from bench import BASE
    from escb200.codec import ESC
    from escb200.spec import CodecSpec
    from escb200.synthetic import synth_audio, synth_state_dict
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 36
    m = ESC(**BASE)
    m.load_state_dict(synth_state_dict(CodecSpec.from_kwargs(**BASE), 0))
    m = m.eval().cuda()
    x = synth_audio(B, 48000, seed=1000).cuda()
    for _ in range(3):
        c, f = m.encode(x, 6)
        m.decode(c, f)
    h = m._handle(torch.device("cuda", 0))
    h.profile_begin()
    c, f = m.encode(x, 6)
    m.decode(c, f)
    h.profile_end()
    sys.exit(0)
env = dict(os.environ, _PL_CHILD="1", ESCB_PROFILE_DUMP="1")
out = subprocess.run([sys.executable, __file__] + sys.argv[1:2], env=env, capture_output=True, text=True)
rows = OrderedDict()
for line in out.stderr.splitlines():
    if line.startswith("escb_launch "):
        _, name, ms, _, fl, _, by, _ = line.split()
        rows.setdefault((name, fl), []).append(float(ms))
if not rows:
    print(out.stderr[-2000:])
sel = sys.argv[2] if len(sys.argv) > 2 else ""
tot = 0.0
for (name, fl), v in rows.items():
    tot += sum(v)
    if sel in name:
        print(f"{name:22s} {float(fl) / 1e9:8.2f} GFLOP  x{len(v):2d}  {1e3 * sum(v) / len(v):7.1f} us each  {sum(v):6.3f} ms")
print(f"total {tot:.2f} ms")
