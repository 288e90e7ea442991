"""World-size-2 gloo test (CPU) of the N>1 host path: contiguous batch shards, one all-gather, clip order kept."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from escb200.parallel import gather_results, shard_batch, shard_bounds


def _fake_codec(x):
    """Deterministic per-clip stand-in for encode+decode (the native library needs a GPU)."""
    codes = (x[:, :18].abs() * 1e4).long().reshape(x.shape[0], 2, 3, 3) % 1024
    audio = x[:, : x.shape[1] - 80] * 0.5
    return codes, audio


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(total, 400, generator=g)
        mine = shard_batch(x, rank, world)
        codes, audio = _fake_codec(mine)
        all_codes, all_audio = gather_results(codes, audio)
        ref_codes, ref_audio = _fake_codec(x)
        ok = torch.equal(all_codes, ref_codes) and torch.equal(all_audio, ref_audio)
        q.put((rank, bool(ok), tuple(all_codes.shape)))
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    assert shard_bounds(288, 0, 8) == (0, 36) and shard_bounds(288, 7, 8) == (252, 288)
    with pytest.raises(ValueError):
        shard_bounds(10, 0, 4)


def test_two_rank_gather_keeps_clip_order():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res)
    assert res[0][2] == (8, 2, 3, 3)


def test_single_process_gather_is_identity():
    c, a = torch.zeros(2, 1, 3, 4, dtype=torch.int64), torch.zeros(2, 10)
    gc, ga = gather_results(c, a)
    assert gc is c and ga is a
