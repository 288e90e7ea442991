"""Small-shape encode + decode + forward(eval) for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):

    compute-sanitizer --tool memcheck python tools/sanitize_target.py

One 1 s clip pair with a ragged width (padded windows, shift masks on every level), all 6 streams, every kernel class
of the library launched at least once; results are compared with a second run to catch nondeterminism too."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-speech-codec_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from escb200.codec import ESC  # noqa: E402
from escb200.spec import CodecSpec  # noqa: E402
from escb200.synthetic import synth_audio, synth_state_dict  # noqa: E402

cfg = dict(codebook_dims=[32, 32, 16, 12, 8, 6])
m = ESC(**cfg)
m.load_state_dict(synth_state_dict(CodecSpec.from_kwargs(**cfg), 0))
m = m.eval().cuda()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
x = synth_audio(B, 16000 + 80 * 4 * 3, seed=9).cuda()
codes, fs = m.encode(x, 6)
audio = m.decode(codes, fs)
out = m(x, None, 6)
torch.cuda.synchronize()
codes2, _ = m.encode(x, 6)
audio2 = m.decode(codes2, fs)
torch.cuda.synchronize()
assert torch.equal(codes, codes2) and torch.equal(audio, audio2) and torch.equal(out["codes"], codes)
h = m._handle(torch.device("cuda", 0))
print(f"sanitize target ok: W={fs[1]} codes {tuple(codes.shape)} launches {h.launch_count()}")
