"""CPU, world_size 2, gloo: the N>1 path — contiguous chunk-aligned shards, no data-path collective, hit gather
and max-over-ranks timing — gives exactly the single-process result.  The scanner here is the oracle (test
infrastructure standing in for the GPU), the host logic under test is keyhunt_b200/sharding.py."""
import os
import random
import sys

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from keyhunt_b200 import sharding

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_range_partitions_exactly():
    for world in (1, 2, 3, 4, 8):
        for chunks in (1, 2, 7, 8, 64, 1000):
            for chunk in (1024, 1 << 20):
                start, n = 0x1234567, chunks * chunk
                cover = []
                for r in range(world):
                    s, m = sharding.shard_range(start, n, world, r, chunk)
                    assert m % chunk == 0
                    cover.append((s, m))
                assert cover[0][0] == start
                for (s0, m0), (s1, _) in zip(cover, cover[1:]):
                    assert s0 + m0 == s1
                assert sum(m for _, m in cover) == n
                assert max(m for _, m in cover) - min(m for _, m in cover) <= chunk
    with pytest.raises(ValueError):
        sharding.shard_range(0, 1000, 2, 0, 1024)


def test_scanned_points_overshoot_rule():
    # -r 1:FFFFFFFF with the default -n 2^32 scans keys 1..2^32
    assert sharding.scanned_points(1, 0xFFFFFFFF, 1 << 32) == 1 << 32
    assert sharding.scanned_points(0x100, 0x100 + 3 * (1 << 20) + 1, 1 << 20) == 4 << 20
    assert sharding.scanned_points(5, 5, 1 << 20) == 0


def test_shard_windows():
    for world in (1, 2, 5, 8):
        for n in (1, 7, 8, 1000):
            parts = [sharding.shard_windows(n, world, r) for r in range(world)]
            assert sum(c for _, c in parts) == n
            pos = 0
            for f, c in parts:
                assert f == pos
                pos += c


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    from _oracle import CRYPTO_BTC, MODE_RMD160, SEARCH_BOTH, Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = Oracle()
    rnd = random.Random(99)
    start, n, chunk = 0x2000000000000000, 6 * 4096, 4096
    idx = [0, n - 1, n // 2, n // 2 - 1] + [rnd.randrange(n) for _ in range(8)]
    recs = []
    for j, i in enumerate(idx):
        x, y = o.pubkey(start + i)
        recs.append(o.hash160_uncomp(x, y) if j % 2 else o.hash160_comp(2 + (y & 1), x))
    t = o.targets_new(b"".join(recs))
    s, m = sharding.shard_range(start, n, world, rank, chunk)
    mine = [(h["key"], h["kind"], h["matched"]) for h in o.scan(t, MODE_RMD160, CRYPTO_BTC, SEARCH_BOTH, s, 1, m, nthreads=1)]
    merged = sharding.gather_hits(dist, mine)
    tmax = sharding.max_over_ranks(dist, 10.0 + rank)
    if rank == 0:
        whole = sorted((h["key"], h["kind"], h["matched"]) for h in o.scan(t, MODE_RMD160, CRYPTO_BTC, SEARCH_BOTH, start, 1, n, nthreads=2))
        q.put((merged == whole, len(merged), tmax, len(mine)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_scan_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, n, tmax, n0 = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same and n == len(set(range(n))) and n >= 10
    assert tmax == 11.0            # max over ranks
    assert 0 < n0 < n              # rank 0 alone does not see every hit


def _worker_bsgs(rank, world, port, q):
    """the host logic of bench.py's strong C4 block / the CLI's `-t N` BSGS: contiguous blocks of 2N-key windows per rank,
    tables built on every rank, found keys gathered — the oracle stands in for the GPU"""
    sys.path.insert(0, HERE)
    from _oracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = Oracle()
    n, k = 1 << 20, 1
    b = o.bsgs_new(n, k, 1)
    start, n_windows = 0x8000000001, 11
    win = 2 * n
    rnd = random.Random(5)
    keys = [start, start + 5 * win + 12345, start + (n_windows - 1) * win + win - 1, start + 3 * win, start + n_windows * win + 7] + \
           [start + rnd.randrange(n_windows * win) for _ in range(3)]
    first, count = sharding.shard_windows(n_windows, world, rank)
    results = []
    for key in keys:
        pub = o.pubkey(key)
        got = o.bsgs_search(b, pub, start + first * win, start + (first + count) * win)[0] if count else None
        merged = sharding.gather_hits(dist, [got] if got is not None else [])
        if rank == 0:
            whole = o.bsgs_search(b, pub, start, start + n_windows * win)[0]
            results.append((key, merged, whole))
    o.bsgs_free(b)
    if rank == 0:
        q.put(results)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bsgs_windows_equal_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_worker_bsgs, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    found = 0
    for key, merged, whole in results:
        assert merged == ([whole] if whole is not None else []), hex(key)
        found += whole is not None
    assert found >= 6          # every key inside the range is found by exactly one rank; the one past the end by nobody
