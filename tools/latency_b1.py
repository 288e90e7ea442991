"""Small-batch latency of encode + decode (one 3 s clip): eager launches vs one CUDA graph replay.
usage: python tools/latency_b1.py [batch=1] [iters=50]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-speech-codec_b200"))
sys.path.insert(0, ROOT)
import torch
from bench import BASE, small_batch_latency
from escb200.codec import ESC
from escb200.spec import CodecSpec
from escb200.synthetic import synth_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda", 0)
m = ESC(**BASE)
m.load_state_dict(synth_state_dict(CodecSpec.from_kwargs(**BASE), 0))
m = m.eval().to(dev)
print(small_batch_latency(m, dev, B, iters))
