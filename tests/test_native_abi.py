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


def _tiling(N, K, role=0):
    out = (ctypes.c_int32 * 8)()
    rc = native.lib().escb_tiling_info(N, K, role, out)
    assert rc == 0, (N, K, role)
    return dict(zip(("BN", "nsub", "ntn", "nkb", "resident", "nmain", "corr", "tmem"), out))


def test_tiling_and_accumulator_split_invariants(monkeypatch):
    """Host logic of the tcgen05 engine (csrc/tc_gemm.cuh choose_tiling / acc_policy), no device needed: every GEMM shape
    of the three shipped configs fits tensor memory with its accumulator split, covers all output columns, never has more
    main accumulators than K blocks, and follows the split policy of DESIGN.md section 2."""
    for k in ("ESCB_ACC", "ESCB_ACC_KMIN", "ESCB_TILE_SERIAL"):
        monkeypatch.delenv(k, raising=False)
    dims = [45, 72, 96, 144, 192, 384]
    shapes = []
    for C in dims:
        shapes += [(C, C, 1), (4 * C, C, 2), (C, 4 * C, 1), (3 * C, C, 0)]             # proj, mlp1, mlp2, unfused qkv
    shapes += [(b, 2 * a, 0) for a, b in zip(dims, dims[1:])]                          # PatchMerge: 2C -> next C
    shapes += [(2 * a, b, 1) for a, b in zip(dims, dims[1:])]                          # PatchSplit: C -> 2 * previous C
    for N, K, role in shapes:
        t = _tiling(N, K, role)
        assert t["BN"] % 16 == 0 and 16 <= t["BN"] <= 208
        assert t["tmem"] == t["BN"] * t["nsub"] * (t["nmain"] + t["corr"]) <= 512, (N, K, t)
        assert t["BN"] * t["nsub"] * t["ntn"] >= N, (N, K, t)
        assert t["nkb"] == (K + 31) // 32 and 1 <= t["nmain"] <= t["nkb"]
        assert t["corr"] == (1 if K > 96 else 0), (N, K, t)                              # corrections split off beyond K = 96
        if K <= 96:
            assert t["nmain"] == 1
        if K >= 1024:
            assert t["nmain"] >= 2, (N, K, t)                                            # long reductions alternate K blocks
    # the round-1 engine is still selectable (A-B runs, variant tests)
    monkeypatch.setenv("ESCB_ACC", "1,0")
    t = _tiling(384, 1536, 1)
    assert (t["nmain"], t["corr"]) == (1, 0)
    monkeypatch.setenv("ESCB_ACC", "4,1")
    t = _tiling(96, 384, 1)
    assert t["corr"] == 1 and t["nmain"] <= 4 and t["tmem"] <= 512
    assert native.lib().escb_tiling_info(0, 32, 0, (ctypes.c_int32 * 8)()) != 0
