"""CPU tests of the wire / on-disk formats and of the host logic of the eval sweep (f1)."""
import numpy as np
import pytest
import torch

from escb200.bitstream import bits_for, load_codes, pack_codes, save_codes, unpack_codes


def test_bitstream_round_trip_and_size(tmp_path):
    g = torch.Generator().manual_seed(0)
    codes = torch.randint(0, 1024, (2, 6, 3, 150), generator=g)
    blob = pack_codes(codes, 1024)
    assert len(blob) == 24 + 2 * 6 * 3 * 150 * 10 // 8            # 9 kbps: 3375 payload bytes per 3 s clip
    back = unpack_codes(blob)
    assert back.dtype == torch.int64 and torch.equal(back, codes)
    n = save_codes(str(tmp_path / "c.escb"), codes)
    assert n == len(blob) and torch.equal(load_codes(str(tmp_path / "c.escb")), codes)


@pytest.mark.parametrize("shape,K", [((1, 1, 3, 1), 1024), ((3, 2, 3, 7), 1024), ((1, 6, 3, 5), 512), ((2, 1, 1, 9), 1000),
                                     ((1, 1, 1, 0), 1024)])
def test_bitstream_ragged_shapes(shape, K):
    g = torch.Generator().manual_seed(1)
    codes = torch.randint(0, K, shape, generator=g) if shape[-1] else torch.zeros(shape, dtype=torch.int64)
    if codes.numel():
        codes.view(-1)[0] = K - 1                                   # extreme values survive
        codes.view(-1)[-1] = 0
    assert bits_for(K) == int(np.ceil(np.log2(K)))
    assert torch.equal(unpack_codes(pack_codes(codes, K)), codes)


def test_bitstream_rejects_bad_input():
    with pytest.raises(IndexError):
        pack_codes(torch.full((1, 1, 3, 4), 1024), 1024)
    with pytest.raises(IndexError):
        pack_codes(torch.full((1, 1, 3, 4), -1), 1024)
    with pytest.raises(ValueError):
        unpack_codes(b"nope")
    blob = pack_codes(torch.zeros((1, 2, 3, 4), dtype=torch.int64))
    with pytest.raises(ValueError):
        unpack_codes(blob[:-1])
    with pytest.raises(ValueError):
        unpack_codes(b"XXXX" + blob[4:])


def _reference_entropy(codes, K):
    """scripts/metrics.py:37-77 of the reference restated with one_hot, as the independent check."""
    B, S, G, T = codes.shape
    ent = {}
    for s in range(S):
        for g in range(G):
            cnt = torch.nn.functional.one_hot(codes[:, s, g], num_classes=K).view(-1, K).sum(0).float()
            p = cnt / float(B * T)
            ent[f"stream_{s}_group_{g + 1}"] = (-torch.sum(p * torch.log2(p + 1e-10))).item()
    rate = round(sum(ent.values()) / (S * G * np.log2(K)), 4)
    return rate, {k: round(v / np.log2(K), 4) for k, v in ent.items()}


def test_entropy_counter_host_path_matches_one_hot_restatement():
    from scripts.metrics import EntropyCounter
    g = torch.Generator().manual_seed(2)
    ec = EntropyCounter(1024, num_streams=6, num_groups=3, device="cpu")
    ec.reset_stats(num_streams=4)
    batches = [torch.randint(0, 1024, (3, 4, 3, 50), generator=g) ** 2 % 1024 for _ in range(3)]
    for c in batches:
        ec.update(c)
    assert ec.total_counts == 3 * 3 * 50
    rate, util = ec.compute_utilization()
    ref_rate, ref_util = _reference_entropy(torch.cat(batches), 1024)
    assert rate == ref_rate and util == ref_util
    assert set(ec.codebook_counts) == {f"stream_{s}_group_{g}" for s in range(4) for g in (1, 2, 3)}
    with pytest.raises(AssertionError, match="size not match"):
        ec.update(torch.zeros(1, 6, 3, 5, dtype=torch.int64))


def test_sisdr_and_eval_set(tmp_path):
    from scripts.metrics import SISDR
    from scripts.utils import EvalSet, save_wav
    x = torch.randn(2, 4000)
    assert (SISDR()(x, x * 0.5) > 60).all()                        # scale invariant
    noisy = x + 0.1 * torch.randn(2, 4000)
    v = SISDR()(x, noisy)
    assert ((v > 15) & (v < 25)).all()                             # ~20 dB
    for i in range(3):
        save_wav(str(tmp_path / f"c{i}.wav"), torch.randn(1, 1000) * 0.1, 16000)
    ds = EvalSet(str(tmp_path))
    assert len(ds) == 3 and tuple(ds[0].shape) == (920,)
