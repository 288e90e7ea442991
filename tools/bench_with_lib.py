"""Run bench.py against another build of the library (A-B experiments; see csrc/Makefile `variant`).
usage: python tools/bench_with_lib.py <path/to/libescb200_xxx.so> [bench.py arguments]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "efficient-speech-codec_b200"))
sys.path.insert(0, ROOT)
from escb200 import native

lib = os.path.abspath(sys.argv[1])
native.library_path = lambda: lib
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[2:]
import bench

bench.main()
