"""RVQCodecs (the reference's rvq+swinT ablation codec): the oracle restatement against the
reference-generated goldens (CPU), and the CUDA path through the C ABI against both (GPU)."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import i64
from escb200.synthetic import synth_audio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RVQ = dict(backbone="transformer", in_dim=2, in_freq=192, h_dims=[45, 72, 96, 144, 192, 384], max_streams=6,
           patch_size=[3, 2], swin_heads=[3, 6, 12, 24, 24], swin_depth=2, window_size=4, mlp_ratio=4.0, overlap=2,
           num_rvqs=6, group_size=3, codebook_size=1024, codebook_dim=8, l2norm=True, win_len=20, hop_len=5, sr=16000)
AUDIO_TOL = 1e-4


@pytest.fixture(scope="module")
def grvq():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_rvq_v1.npz"))


def maxabs(a, b):
    return float((torch.as_tensor(np.asarray(a)).double() - torch.as_tensor(np.asarray(b)).double()).abs().max())


def make_rvq_oracle(seed):
    from escb200.spec import CodecSpec
    from escb200.synthetic import synth_state_dict
    from oracle.esc_oracle import RvqOracle
    sd = synth_state_dict(CodecSpec.from_rvq_kwargs(**RVQ), seed)
    return RvqOracle(RVQ, sd), sd


def make_rvq_native(seed):
    from esc.models import make_model
    from escb200.spec import CodecSpec
    from escb200.synthetic import synth_state_dict
    m = make_model(RVQ, "rvq+swinT")
    m.load_state_dict(synth_state_dict(CodecSpec.from_rvq_kwargs(**RVQ), seed), strict=True)
    return m.eval().cuda()


# ---------------------------------------------------------------------------------------------- CPU: oracle vs reference
def test_rvq_state_dict_is_the_reference_manifest():
    from esc import RVQCodecs
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_manifest_rvq.json")))["rvq_swinT"]
    sd = RVQCodecs(**RVQ).state_dict()
    mine = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()}
    assert mine == {k: [s, d] for k, s, d in ref}
    with pytest.raises(TypeError):
        RVQCodecs(codebook_dims=[8] * 6)                  # ESC's kwarg, not RVQCodecs' (codecs.py:98-119)
    with pytest.raises(TypeError):
        RVQCodecs(**dict(RVQ, codebook_dim=[8] * 6))      # configs/ablations/9kbps_rvq_conv.yaml: fails in the reference too


def test_rvq_oracle_matches_reference_goldens(grvq):
    o, _ = make_rvq_oracle(8)
    x = synth_audio(2, 48000, seed=9)[:1]
    codes, fs = o.encode(x, 6)
    assert fs == (2, 300)
    assert torch.equal(codes, i64(grvq["E_codes"])[:1])
    assert maxabs(o.decode(codes, fs), grvq["E_audio"][:1]) <= 1e-5
    c3, _ = o.encode(x, 3)
    assert torch.equal(c3, codes[:, :3])
    assert maxabs(o.decode(c3, fs)[0, :4000], grvq["E_audio_s3_head"]) <= 1e-5
    o2, _ = make_rvq_oracle(10)
    xf = torch.from_numpy(grvq["F_x"])
    cf, fsf = o2.encode(xf, 4)
    assert fsf == (2, 10) and torch.equal(cf, i64(grvq["F_codes"]))
    assert maxabs(o2.decode(cf, fsf), grvq["F_audio"]) <= 1e-5
    fo = o2.forward(xf, None, 2)
    assert torch.equal(fo["codes"], i64(grvq["F_fwd2_codes"]))
    assert maxabs(fo["cm_loss"], grvq["F_fwd2_cm_loss"]) <= 1e-6
    assert maxabs(fo["recon_audio"], grvq["F_fwd2_audio"]) <= 1e-5


# ---------------------------------------------------------------------------------------------- GPU: CUDA path
@pytest.mark.gpu
def test_rvq_golden_e_all_bitrates(grvq):
    m = make_rvq_native(8)
    x = synth_audio(2, 48000, seed=9).cuda()
    codes, fs = m.encode(x, 6)
    assert fs == (2, 300) and codes.dtype == torch.int64
    assert torch.equal(codes.cpu(), i64(grvq["E_codes"]))
    audio = m.decode(codes, fs)
    assert tuple(audio.shape) == (2, 47920)
    assert maxabs(audio.cpu(), grvq["E_audio"]) <= AUDIO_TOL
    for s in range(1, 6):
        cs, _ = m.encode(x[:1], s)
        assert torch.equal(cs.cpu(), i64(grvq["E_codes"])[:1, :s])
        assert maxabs(m.decode(cs, fs)[0, :4000].cpu(), grvq[f"E_audio_s{s}_head"]) <= AUDIO_TOL
    out = m(x[:1, :-80], None, 6)
    assert torch.equal(out["codes"].cpu(), i64(grvq["E_fwd_codes"]))
    assert maxabs(out["cm_loss"].cpu(), grvq["E_fwd_cm_loss"]) <= 1e-4
    assert maxabs(out["cb_loss"].cpu(), grvq["E_fwd_cb_loss"]) <= 1e-4
    assert maxabs(out["recon_audio"][0, :4000].cpu(), grvq["E_fwd_audio_head"]) <= AUDIO_TOL
    # forward(eval) == decode(encode(x)), like the reference
    ce, fse = m.encode(x[:1, :-80], 6)
    assert torch.equal(out["codes"], ce) and torch.equal(out["recon_audio"], m.decode(ce, fse))


@pytest.mark.gpu
def test_rvq_golden_f_ragged_and_errors(grvq):
    from escb200 import native
    m = make_rvq_native(10)
    x = torch.from_numpy(grvq["F_x"]).cuda()
    codes, fs = m.encode(x, 4)
    assert fs == (2, 10) and torch.equal(codes.cpu(), i64(grvq["F_codes"]))
    assert maxabs(m.decode(codes, fs).cpu(), grvq["F_audio"]) <= AUDIO_TOL
    fo = m(x, None, 2)
    assert torch.equal(fo["codes"].cpu(), i64(grvq["F_fwd2_codes"]))
    assert maxabs(fo["cm_loss"].cpu(), grvq["F_fwd2_cm_loss"]) <= 1e-4
    assert maxabs(fo["recon_audio"].cpu(), grvq["F_fwd2_audio"]) <= AUDIO_TOL
    assert maxabs(fo["recon_feat"].cpu(), grvq["F_fwd2_recon_feat"]) <= 2e-4
    # host tensors take the *_host entry points
    ch, _ = m.encode(x.cpu(), 4)
    assert torch.equal(ch, codes.cpu())
    assert torch.equal(m.decode(ch, fs), m.decode(codes, fs).cpu())
    with pytest.raises(native.NativeError):
        m.encode(x, 7)
    with pytest.raises(IndexError):
        bad = codes.cpu().clone()
        bad[0, 0, 0, 0] = 4096
        m.decode(bad, fs)


@pytest.mark.gpu
def test_rvq_batch_matches_oracle():
    """12 clips of 1 s against the oracle on two of them, batch invariance."""
    m = make_rvq_native(12)
    o, _ = make_rvq_oracle(12)
    x = synth_audio(12, 16000, seed=13).cuda()
    codes, fs = m.encode(x, 6)
    audio = m.decode(codes, fs)
    for b in (0, 11):
        co, _ = o.encode(x[b:b + 1].cpu(), 6)
        assert torch.equal(co[0], codes[b].cpu())
        assert maxabs(o.decode(co, fs)[0], audio[b].cpu()) <= AUDIO_TOL
        c1, _ = m.encode(x[b:b + 1], 6)
        assert torch.equal(c1[0], codes[b])
