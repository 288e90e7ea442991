"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/escb200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from escb200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "escb200.h")).read()
    text = re.sub(r"#ifdef ESCB_TC_TRACE.*?#endif", "", text, flags=re.S)      # debug-build-only exports
    return re.findall(r"^ESCB_API [\w\* ]+?\b(escb_\w+)\(", text, flags=re.M)


def test_header_and_binding_agree():
    syms = header_symbols()
    assert len(syms) == len(set(syms)) >= 26
    assert sorted(syms) == sorted(native.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(native.library_path())
    for s in header_symbols():
        assert hasattr(lib, s), s
    assert native.lib().escb_abi_version() == native.ESCB_ABI_VERSION


def test_config_struct_layout_matches_header():
    # 6 + 8 + 8 + 6 + 8 + 2 int32 fields
    assert ctypes.sizeof(native.EscbConfig) == 4 * (6 + 8 + 8 + 6 + 8 + 2)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from escb200.codec import ESC
    from escb200.spec import CodecSpec
    with pytest.raises(native.NativeError) as ei:
        native.Handle(CodecSpec.from_kwargs())
    assert ei.value.code == -2          # ESCB_ENODEV
    m = ESC().eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.encode(torch.zeros(1, 16000), 6)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.decode(torch.zeros(1, 6, 3, 50, dtype=torch.int64), (2, 100))


def test_state_dict_is_the_reference_manifest():
    import json
    from helpers import BASE, LARGE
    from escb200.codec import ESC
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_manifest.json")))
    for name, cfg in (("base", BASE), ("large", LARGE)):
        sd = ESC(**cfg).state_dict()
        mine = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()}
        theirs = {k: [s, d] for k, s, d in ref[name]}
        assert mine == theirs


def test_ctor_rejects_what_the_reference_rejects():
    from escb200.codec import ESC, make_model
    with pytest.raises(TypeError):
        ESC(codebook_dim=8)                       # configs/ablations/9kbps_csvq_conv.yaml:21 vs codecs.py:16
    with pytest.raises(NotImplementedError):
        make_model({}, "dac+base")                # not one of the reference's four model names
    assert type(make_model({}, "rvq+swinT")).__name__ == "RVQCodecs"
    m = make_model(dict(swin_depth=4))
    assert m.max_streams == 6 and m.max_bps == 9.0
    with pytest.raises(AssertionError, match="multiple of overlap"):
        m.time_patches(16000 + 160)               # W = 101
