"""N>1 host logic on CPU: world_size-2 gloo group, contiguous batch split, no data-path collective, max-over-ranks timing.
The per-shard compute is done by the oracle here (this is a CPU test of the sharding plumbing, not of the kernels)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from oracle import pyoracle as O
    from poulpy_b200.sharding import gather_shards, max_over_ranks, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, k = 64, 18
    rng = np.random.default_rng(1)  # same inputs on every rank
    a = rng.integers(-(1 << 17), 1 << 17, size=(total, 3, 2, n), dtype=np.int64)
    mat = rng.integers(-(1 << 17), 1 << 17, size=(3, 1, 4, 2, n), dtype=np.int64)
    om = O.OracleModule(n, O.NTT120)
    pm = om.vmp_pmat_alloc(3, 1, 2, 4)
    om.vmp_prepare(pm, mat)  # "key replicated once per rank"
    lo, hi = shard_range(total, rank, world)
    mine = np.zeros((hi - lo, 3, 2, n), dtype=np.int64)
    if hi > lo:
        om.glwe_keyswitch_batch(mine, k, np.ascontiguousarray(a[lo:hi]), k, pm, k, 1, threads=1)
    full = gather_shards(mine, total)
    t = max_over_ranks(float(rank + 1))
    if rank == 0:
        want = np.zeros_like(a)
        om.glwe_keyswitch_batch(want, k, a, k, pm, k, 1, threads=1)
        q.put((bool(np.array_equal(full, want)), t, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [7, 8])
def test_two_rank_sharding(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, tmax, rng0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and tmax == 2.0 and rng0 == (0, (total + 1) // 2)


def test_shard_range_partitions():
    from poulpy_b200.sharding import shard_range

    for total in (0, 1, 5, 8, 4096, 4097):
        for world in (1, 2, 4, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1


def test_broadcast_prepared_is_a_noop_without_a_group():
    """Single-process runs (N = 1, the tests, smoke) never initialise torch.distributed: the key replication helper must not touch its
    arguments then (its NCCL leg is exercised by the 2- and 8-GPU bench runs, profiles/r1_bench_v10_n2.json / r1_bench_v11_n8.json)."""
    from poulpy_b200.sharding import _DeviceBytes, broadcast_prepared
    assert broadcast_prepared(None, None) is None
    view = _DeviceBytes(0x7F0000000000, 4096).__cuda_array_interface__
    assert view["shape"] == (4096,) and view["typestr"] == "|u1" and view["data"] == (0x7F0000000000, False)


def _replicate_worker(rank, world, port, fail_first, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from poulpy_b200.sharding import max_over_ranks, replicate_prepared

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    key = torch.zeros(4, dtype=torch.int64)
    calls = {"prepare": 0, "broadcast": 0}

    def prepare():
        calls["prepare"] += 1
        if fail_first and rank == 0 and calls["prepare"] == 1:
            raise RuntimeError("injected prepare failure on the source rank")
        key[:] = torch.arange(4) + 10

    def broadcast():
        calls["broadcast"] += 1
        dist.broadcast(key, src=0)

    how = replicate_prepared(prepare, broadcast)
    t = max_over_ranks(float(rank))  # default device under gloo: the host
    q.put((rank, how, calls["prepare"], calls["broadcast"], key.tolist(), t))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_first", [False, True])
def test_replicate_prepared_takes_one_branch_on_every_rank(fail_first):
    """Key replication: rank 0 prepares and broadcasts; when rank 0's prepare raises, NO rank enters the broadcast (nobody hangs) and every
    rank prepares locally -- the status flag makes the branch symmetric."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_replicate_worker, args=(r, 2, port, fail_first, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, how, n_prep, n_bc, key, t in got:
        assert key == [10, 11, 12, 13] and t == 1.0
        if fail_first:
            assert how == "local" and n_bc == 0 and n_prep == (2 if rank == 0 else 1)
        else:
            assert how == "broadcast" and n_bc == 1 and n_prep == (1 if rank == 0 else 0)
