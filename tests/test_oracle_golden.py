"""Pin the CPU oracle (oracle/esc_oracle.py) against outputs of the REAL reference.

tests/golden/golden_v1.npz was produced by tests/golden/make_golden.py, which imports the reference
read-only from /root/reference.  The reference ships no golden vectors of its own, so
these reference-generated fixtures are what pins parity.  On the machine that generated them the oracle
is bit-identical; elsewhere BLAS kernels may differ, hence codes exact + float tolerance 2e-5.
"""
import numpy as np
import pytest
import torch

from helpers import ADV, BASE, LARGE, i64, make_oracle
from escb200.synthetic import synth_audio

TOL = 2e-5


def close(a, b, tol=TOL):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max()) <= tol


def test_case_a_base_3s_all_bitrates(golden):
    o, _ = make_oracle(BASE, 0)
    x = synth_audio(2, 48000, seed=1)
    codes, fs = o.encode(x, 6)
    assert fs == (2, 300)
    assert codes.dtype == torch.int64 and tuple(codes.shape) == (2, 6, 3, 150)
    assert torch.equal(codes, i64(golden["A_codes"]))
    audio = o.decode(codes, fs)
    assert tuple(audio.shape) == (2, 47920)
    assert close(audio, golden["A_audio"])
    for s in range(1, 6):
        cs, _ = o.encode(x[:1], s)
        assert torch.equal(cs, codes[:1, :s])
        au = o.decode(cs, fs)
        assert close(au[0, :4000], golden[f"A_audio_s{s}_head"])
        sums = golden[f"A_audio_s{s}_sums"]
        assert abs(au.double().abs().sum().item() - sums[1]) < 1e-2


def test_case_a_forward_eval_matches_encode_decode(golden):
    o, _ = make_oracle(BASE, 0)
    x = synth_audio(2, 48000, seed=1)[:1, :-80]
    out = o.forward(x, None, 6)
    assert torch.equal(out["codes"], i64(golden["A_fwd_codes"]))
    assert close(out["cm_loss"], golden["A_fwd_cm_loss"], 1e-4)
    assert close(out["cb_loss"], golden["A_fwd_cb_loss"], 1e-4)
    assert close(out["recon_audio"][0, :4000], golden["A_fwd_audio_head"])
    assert out["raw_feat"].shape[-1] == 600 and out["recon_feat"].shape[-1] == 600


def test_case_b_large(golden):
    o, _ = make_oracle(LARGE, 2)
    x = synth_audio(1, 48000, seed=3)
    codes, fs = o.encode(x, 6)
    assert torch.equal(codes, i64(golden["B_codes"]))
    assert close(o.decode(codes, fs), golden["B_audio"])


def test_case_c_adv_dims_1s(golden):
    o, _ = make_oracle(ADV, 4)
    x = synth_audio(2, 16000, seed=5)
    codes, fs = o.encode(x, 6)
    assert fs == (2, 100)
    assert torch.equal(codes, i64(golden["C_codes"]))
    assert close(o.decode(codes, fs), golden["C_audio"])


def test_case_d_ragged_taps(golden):
    """W=10 is not a multiple of the 4x4 window: exercises zero padding after norm1 and the cyclic shift."""
    o, _ = make_oracle(BASE, 6)
    o.record_taps = True
    x = torch.from_numpy(golden["D_x"])
    codes, fs = o.encode(x, 6)
    enc_taps = dict(o.taps)
    o.taps = {}
    audio = o.decode(codes, fs)
    assert fs == (2, 10)
    assert torch.equal(codes, i64(golden["D_codes"]))
    assert close(audio, golden["D_audio"])
    assert close(enc_taps["stft"], golden["D_stft"])
    checked = 0
    for k, v in list(enc_taps.items()) + list(o.taps.items()):
        gk = "D_tap_" + k
        if gk in golden.files:
            assert close(v, golden[gk]), k
            checked += 1
    assert checked >= 14


@pytest.mark.parametrize("s", [3, 6])
def test_case_d_forward_eval(golden, s):
    o, _ = make_oracle(BASE, 6)
    out = o.forward(torch.from_numpy(golden["D_x"]), None, s)
    assert torch.equal(out["codes"], i64(golden[f"D_fwd{s}_codes"]))
    assert out["codes"].shape[1] == s
    assert close(out["cm_loss"], golden[f"D_fwd{s}_cm_loss"], 1e-4)
    assert close(out["recon_audio"], golden[f"D_fwd{s}_audio"])
    assert close(out["recon_feat"], golden[f"D_fwd{s}_recon_feat"])


def test_pvq_layer_and_ties(golden):
    from oracle.esc_oracle import OracleConfig, codebook_argmin, pvq_decode, pvq_encode
    o, sd = make_oracle(BASE, 6)
    cfg = OracleConfig(**BASE)
    z = torch.from_numpy(golden["D_pvq3_in"])
    codes = pvq_encode(sd, "quantizers.3", z, 8, cfg)
    assert torch.equal(codes, i64(golden["D_pvq3_codes"]))
    assert close(pvq_decode(sd, "quantizers.3", codes, 8, cfg), golden["D_pvq3_dec"], 1e-6)
    # duplicated rows and rows that coincide after L2 normalisation: the first (lowest) index wins
    t = codebook_argmin(torch.from_numpy(golden["D_tie_z"]), torch.from_numpy(golden["D_tie_table"]), True)
    assert torch.equal(t, i64(golden["D_tie_codes"]))
    assert t[0, 0].item() == 5 and t[0, 1].item() == 3


def test_manifest_equals_reference_state_dict():
    """Key / shape / dtype list of our spec == the reference ESC.state_dict() (captured by make_golden.py)."""
    import json, os
    from escb200.spec import CodecSpec
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "state_dict_manifest.json")))
    for name, cfg in (("base", BASE), ("large", LARGE)):
        mine = [[e.key, list(e.shape), e.dtype] for e in CodecSpec.from_kwargs(**cfg).manifest()]
        assert mine == ref[name]
    assert len(ref["base"]) == 430
