"""Multi-GPU plumbing on CPU: a batch splits into contiguous per-rank ranges with no data-path
collective; world_size-2 gloo run where each rank processes its shard (the host
simulator stands in for the device) and rank 0 checks the concatenation against a single-rank run."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from libgoldilocks_b200.engine import shard_messages, shard_range, shard_ranges


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 8, 1000, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 4, 8):
            r = shard_ranges(n, world)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_shard_messages_rebases_offsets():
    off = np.array([0, 3, 3, 10, 14], dtype=np.uint64)
    lo, hi, sub = shard_messages(off, 1, 3)
    assert (lo, hi) == (3, 10) and list(sub) == [0, 0, 7]


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import util
    sim = util.hostsim_lib()
    n = 24
    sig, pk, msgs, kinds = util.verify_corpus(sim, "shard", n, corrupt_every=4)
    lo, hi = shard_range(n, rank, world)
    st = sim.ed448_verify(sig[lo:hi], pk[lo:hi], msgs[lo:hi])
    st_rlc, _ = sim.ed448_verify_rlc(sig[lo:hi], pk[lo:hi], msgs[lo:hi])   # per-shard batch equation (own weights), same statuses
    assert (st_rlc == st).all()
    ok = np.flatnonzero(kinds[lo:hi] == 0) + lo                            # the shard's untouched entries alone: the equation decides
    st_ok, fast = sim.ed448_verify_rlc(sig[ok], pk[ok], [msgs[i] for i in ok])
    assert fast == 1 and (st_ok == -1).all()
    u = util.stream_bytes("shard/u", n * 56).reshape(n, 56)
    k = util.stream_bytes("shard/k", n * 56).reshape(n, 56)
    xo, _ = sim.x448(u[lo:hi], k[lo:hi])
    np.save(os.path.join(tmp, "st%d.npy" % rank), st)
    np.save(os.path.join(tmp, "x%d.npy" % rank), xo)
    dist.barrier()   # the only collective: completion, never data
    if rank == 0:
        full = sim.ed448_verify(sig, pk, msgs)
        got = np.concatenate([np.load(os.path.join(tmp, "st%d.npy" % r)) for r in range(world)])
        assert (got == full).all()
        assert (got[kinds == 0] == -1).all()
        xs = np.concatenate([np.load(os.path.join(tmp, "x%d.npy" % r)) for r in range(world)])
        assert (xs == sim.x448(u, k)[0]).all()
    dist.destroy_process_group()


def test_two_rank_sharded_verify_gloo(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)


# ---- the library's own partition of ONE batch over a device set (goldilocks_b200_set_devices, csrc/shard.h) ----

def _lib():
    import libgoldilocks_b200 as g
    return g.load()


def test_library_shard_plan_heavy_is_one_contiguous_range_per_device():
    lib = _lib()
    for n in (4096, 4097, 1 << 20, (1 << 20) + 5, (1 << 24) + 1):
        for ndev in (1, 2, 3, 4, 8):
            pl = lib.shard_plan(n, ndev)
            assert len(pl) == min(ndev, n // 2048)
            assert pl[0][0] == 0 and pl[-1][1] == n
            assert all(pl[i][1] == pl[i + 1][0] for i in range(len(pl) - 1))
            sizes = [b - a for a, b, _, _ in pl]
            assert max(sizes) - min(sizes) <= 1
            assert [s for _, _, s, _ in pl] == list(range(len(pl))) and all(l == 0 for _, _, _, l in pl)
    assert lib.shard_plan(0, 8) == []
    assert lib.shard_plan(100, 8) == [(0, 100, 0, 0)]          # too small to cut


def test_library_shard_plan_light_pipelines_chunks_over_lanes():
    lib = _lib()
    for n, ndev, bpe in (((1 << 24), 8, 312), ((1 << 20) + 3, 2, 768), (1 << 21, 1, 316), (10000, 4, 168)):
        pl = lib.shard_plan(n, ndev, bpe, pipelined=True)
        assert pl[0][0] == 0 and pl[-1][1] == n
        assert all(pl[i][1] == pl[i + 1][0] for i in range(len(pl) - 1))       # contiguous, in order
        assert all(pl[i][2] <= pl[i + 1][2] for i in range(len(pl) - 1))       # a device owns one contiguous region
        per_dev = {}
        for a, b, s, l in pl:
            per_dev.setdefault(s, []).append((b - a, l))
        assert len(per_dev) == min(ndev, n // 2048)
        for s, chunks in per_dev.items():
            assert [l for _, l in chunks] == [i % 3 for i in range(len(chunks))]   # lanes dealt round-robin
            sizes = [c for c, _ in chunks]
            assert max(sizes) - min(sizes) <= 1
            assert len(chunks) == 1 or max(sizes) * bpe <= (24 << 20) + bpe * 4096
        totals = [sum(c for c, _ in v) for v in per_dev.values()]
        assert max(totals) - min(totals) <= 1


def test_set_devices_fails_loudly_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib()
    with pytest.raises(RuntimeError):
        lib.set_devices([0])
    assert lib.get_devices() == []
