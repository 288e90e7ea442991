"""Per-launch times over 40 profiled encode+decode steps: reports kernel classes with launches slower than 1.5x their median
(rare barrier stalls would show up here).  usage: python tools/rare_stalls.py  (GPU box)"""
import collections
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("_CHILD"):
    sys.path.insert(0, os.path.join(ROOT, "efficient-speech-codec_b200")); sys.path.insert(0, ROOT)
    import torch
    from bench import BASE
    from escb200.codec import ESC
    from escb200.spec import CodecSpec
    from escb200.synthetic import synth_audio, synth_state_dict
    m = ESC(**BASE); m.load_state_dict(synth_state_dict(CodecSpec.from_kwargs(**BASE), 0)); m = m.eval().cuda()
    x = synth_audio(36, 48000, seed=1000).cuda()
    for _ in range(3):
        c, f = m.encode(x, 6); m.decode(c, f)
    h = m._handle(torch.device("cuda", 0))
    for it in range(40):
        h.profile_begin(); c, f = m.encode(x, 6); m.decode(c, f); h.profile_end()
    sys.exit(0)
env = dict(os.environ, _CHILD="1", ESCB_PROFILE_DUMP="1")
out = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
rows = collections.defaultdict(list)
for line in out.stderr.splitlines():
    if line.startswith("escb_launch "):
        _, name, ms, _, fl, _, by, _ = line.split()
        rows[(name, fl)].append(float(ms))
for (name, fl), v in rows.items():
    v2 = sorted(v)
    med = v2[len(v2)//2]
    if v2[-1] > 1.5 * med:
        print(f"{name:22s} {float(fl)/1e9:7.2f} GF n={len(v)} median {med*1e3:7.1f} us max {v2[-1]*1e3:8.1f} us  >1.5x: {sum(1 for t in v if t > 1.5*med)}")
print("done", sum(len(v) for v in rows.values()), "launches")
