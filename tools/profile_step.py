"""One warm-up + N measured encode+decode steps at the bench workload, with no torch kernels in between —
the command ncu wraps (see profiles/README.md).  usage: python tools/profile_step.py [batch] [steps] [config]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-speech-codec_b200"))
sys.path.insert(0, ROOT)
import torch
from bench import BASE, LARGE
from escb200.codec import ESC
from escb200.spec import CodecSpec
from escb200.synthetic import synth_audio, synth_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 36
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = LARGE if (len(sys.argv) > 3 and sys.argv[3] == "large") else BASE
m = ESC(**cfg)
m.load_state_dict(synth_state_dict(CodecSpec.from_kwargs(**cfg), 0))
m = m.eval().cuda()
x = synth_audio(B, 48000, seed=1000).cuda()
for _ in range(1 + steps):
    codes, fs = m.encode(x, 6)
    audio = m.decode(codes, fs)
torch.cuda.synchronize()
print("launches", m._handle(torch.device("cuda", 0)).launch_count())
