"""GPU parity tests for the BSGS path (through the C ABI) against the CPU oracle: byte-identical
3-tier bloom shards and bP table, identical found keys."""
import hashlib
import json
import os
import random
import time

import pytest

from _oracle import N_ORDER, P_FIELD

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _canon_table(raw):
    """(6-byte key, index) pairs sorted with ties by index — the reference sort is not stable"""
    ents = [(raw[i:i + 6], int.from_bytes(raw[i + 8:i + 16], "little")) for i in range(0, len(raw), 16)]
    return sorted(ents)


@pytest.mark.parametrize("n,k", [(1 << 20, 1), (1 << 22, 2), (1 << 24, 4), (1 << 26, 3)])
def test_bsgs_build_is_byte_identical(kh, oracle, n, k):
    kh.bsgs_build(n, k)
    d = kh.bsgs_describe()
    b = oracle.bsgs_new(n, k)
    try:
        p = oracle.bsgs_params(b)
        assert (d.n, d.m, d.m2, d.m3, d.aux) == (p["n"], p["m"], p["m2"], p["m3"], p["aux"])
        for tier in (1, 2, 3):
            assert d.tier[tier - 1].as_dict() == oracle.bloom_desc(oracle.bsgs_bloom(b, tier, 0))
            for shard in range(256):
                assert kh.bsgs_export(tier, shard) == oracle.bloom_bytes(oracle.bsgs_bloom(b, tier, shard)), (tier, shard)
        got = kh.bsgs_export(0)
        assert _canon_table(got) == _canon_table(oracle.bsgs_table(b))
        # already canonical on the device: ascending (key, index)
        assert _canon_table(got) == [(got[i:i + 6], int.from_bytes(got[i + 8:i + 16], "little")) for i in range(0, len(got), 16)]
    finally:
        oracle.bsgs_free(b)


@pytest.mark.parametrize("n,k", [(1 << 20, 1), (1 << 24, 4), (1 << 26, 3)])
def test_bsgs_binned_build_equals_direct(kh, n, k):
    """option "bsgs_binned_build": baby points binned by the first 12 bits of X and applied bucket by bucket (the way big tables
    are built) give the same bytes as setting every bit directly: digests of every tier, of the bP table and of the prefix bitmap,
    and the raw bytes of sampled shards."""
    seen = {}
    try:
        for mode, slice_kb in ((0, None), (2, None), (2, "1")):      # a 1 KB slice: the shards are applied in many passes (as at -k 4096)
            kh.set_option("bsgs_binned_build", mode)
            if slice_kb:
                os.environ["KH_BABY_SLICE_KB"] = slice_kb
            kh.bsgs_build(n, k)
            seen[(mode, slice_kb)] = [kh.bsgs_digest(t) for t in range(5)] + [hashlib.sha256(kh.bsgs_export(1, s)).hexdigest() for s in (0, 17, 255)]
    finally:
        os.environ.pop("KH_BABY_SLICE_KB", None)
        kh.set_option("bsgs_binned_build", 1)
    assert seen[(0, None)] == seen[(2, None)] == seen[(2, "1")]
    assert len(set(seen[(0, None)][:5])) == 5


def test_bsgs_search_finds_planted_keys(kh, oracle):
    n, k = 1 << 24, 2
    kh.bsgs_build(n, k)
    b = oracle.bsgs_new(n, k)
    try:
        p = oracle.bsgs_params(b)
        m, m2, m3 = p["m"], p["m2"], p["m3"]
        start = 0x8000000000
        end = start + 40 * 2 * p["n"]
        rnd = random.Random(3)
        keys = [start, start + 1, start + m, start + m - 1, start + m + 1, start + 2 * m, end - 1,
                start + 5 * 2 * m + (2 * 3 + 1) * m2,            # exact tier-2 centre
                start + 9 * 2 * m + 4 * 2 * m2 + (2 * 7 + 1) * m3,  # exact tier-3 centre (special case keyhunt.cpp:5238)
                ] + [rnd.randrange(start, end) for _ in range(24)]
        for key in keys:
            pub = oracle.pubkey(key)
            want, _, _ = oracle.bsgs_search(b, pub, start, end)
            got = kh.bsgs_search(pub, start, end)
            assert got == want, hex(key)
            if want is not None:
                assert want == key
        # a key outside the range is not found (and the walk ends)
        pub = oracle.pubkey(end + 10 * p["n"])
        assert kh.bsgs_search(pub, start, end) == oracle.bsgs_search(b, pub, start, end)[0]
    finally:
        oracle.bsgs_free(b)


def test_bsgs_reference_fixture_puzzles(kh, oracle):
    """tests/1to63_65.txt lines 21..32 (puzzle keys) with `-n 0x400000 -k 2 -r 100000:10000000000`; the
    found keys are the ones the unmodified reference binary printed (tests/golden/README.md)"""
    kh.bsgs_build(1 << 22, 2)
    keys = [0x1ba534, 0x2de40f, 0x556e52, 0xdc2a04, 0x1fa5ee5, 0x340326e, 0x6ac3875, 0xd916ce8, 0x17e2551e, 0x3d94cd64,
            0x7d4fe747, 0xb862a62e]
    for key in keys:
        assert kh.bsgs_search(oracle.pubkey(key), 0x100000, 0x10000000000) == key


def test_c4_full_size_properties(kh, oracle):
    """BASELINE config 4 at its real size (bsgs -k 512, n = 2^44: m = 2^31 baby points, 7.7 GB tier-1 bloom) through
    size-independent properties: the reference's sizes, membership of sampled baby points in every tier they belong
    to, bP table = sorted permutation with the right keys, planted keys found in the first / a middle / the last window."""
    kh.bsgs_build(1 << 44, 512)
    d = kh.bsgs_describe()
    assert (d.m, d.m2, d.m3, d.aux) == (1 << 31, 1 << 26, 1 << 21, 8192)
    assert [d.tier[i].bytes for i in range(3)] == [30151987, 942250, 35944]          # SURVEY App. A.4 (reference-probed)
    rnd = random.Random(44)
    shards = {}

    def member(tier, x):
        xb = x.to_bytes(32, "big")
        if (tier, xb[0]) not in shards:
            shards[(tier, xb[0])] = kh.bsgs_export(tier, xb[0])
        bf, desc = shards[(tier, xb[0])], d.tier[tier - 1]
        a = oracle.xxh64(xb, 0x59f2815b16f81798)
        b = oracle.xxh64(xb, a)
        return all((bf[(((a + b * i) & (2**64 - 1)) % desc.bits) >> 3] >> ((((a + b * i) & (2**64 - 1)) % desc.bits) & 7)) & 1
                   for i in range(desc.hashes))

    for tier, top in ((3, d.m3), (2, d.m2), (1, d.m)):
        assert all(member(tier, oracle.pubkey(j)[0]) for j in [1, top] + [rnd.randrange(1, top + 1) for _ in range(6)])
    shards.clear()
    assert not member(3, oracle.pubkey(d.m3 + 1)[0]) or not member(3, oracle.pubkey(d.m3 + 2)[0])   # beyond the tier: not inserted
    shards.clear()
    tab = kh.bsgs_export(0)
    ents = [(tab[i:i + 6], int.from_bytes(tab[i + 8:i + 16], "little")) for i in range(0, len(tab), 16)]
    assert all(ents[i] <= ents[i + 1] for i in range(len(ents) - 1))
    assert sorted(e[1] for e in ents) == list(range(d.m3))
    for i in rnd.sample(range(d.m3), 12):
        assert ents[i][0] == oracle.pubkey(ents[i][1] + 1)[0].to_bytes(32, "big")[16:22]
    lo, hi = 1 << 64, 1 << 65
    for key in (lo + 5, lo + 37 * (1 << 45) + rnd.randrange(1 << 45), hi - 12345):
        assert kh.bsgs_search(oracle.pubkey(key), lo, hi) == key
    st = kh.stats()
    assert st["tier1_positives"] > 0
    # the tables above were built binned (the default for >= 2^26 baby steps); a direct build holds the same bytes in all
    # 7.7 GB of tier 1 and all 64 GB of the prefix bitmap (device-side digests), and takes several times longer
    binned = [kh.bsgs_digest(t) for t in range(5)]
    kh.stats(reset=True)
    t0 = time.time(); kh.bsgs_build(1 << 44, 512); t_binned = time.time() - t0
    ms_binned = kh.stats(reset=True)["walk_ms"]
    assert [kh.bsgs_digest(t) for t in range(5)] == binned
    try:
        kh.set_option("bsgs_binned_build", 0)
        t0 = time.time(); kh.bsgs_build(1 << 44, 512); t_direct = time.time() - t0
        ms_direct = kh.stats(reset=True)["walk_ms"]
        assert [kh.bsgs_digest(t) for t in range(5)] == binned
    finally:
        kh.set_option("bsgs_binned_build", 1)
    try:        # evidence for profiles/ (the GPU box merges gpurun_out/ back)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump({"what": "kh_bsgs_build(n = 2^44, k = 512): 2^31 baby steps into the 7.7 GB tier-1 bloom + 64 GB prefix bitmap, seconds",
                   "binned_wall_s": t_binned, "direct_wall_s": t_direct, "binned_walk_ms": ms_binned, "direct_walk_ms": ms_direct,
                   "baby_steps_per_s_binned": d.m / (ms_binned * 1e-3), "baby_steps_per_s_direct": d.m / (ms_direct * 1e-3),
                   "digests_equal": True},
                  open(os.path.join(ROOT, "gpurun_out", "bsgs_build_binned.json"), "w"))
    except OSError:
        pass


def test_bsgs_server_variant_base_check(kh, oracle):
    """option "bsgs_base_check" = the reference SERVER's loop (bsgsd.cpp:2544): a key equal to the base of a 2N window
    is reported; without it (keyhunt.cpp's loop) the same key stays unseen.  Both against the oracle's two variants."""
    n, k = 1 << 24, 4
    kh.bsgs_build(n, k)
    b = oracle.bsgs_new(n, k)
    try:
        m = oracle.bsgs_params(b)["m"]
        start = 0x8000000001
        end = start + 6 * 2 * n
        keys = [start, start + 2 * n, start + 5 * 2 * n, start + 6 * 2 * n, start + 1, start + m, start + 2 * m, start + 2 * m + 1,
                start + 2 * n - 1, start + 2 * n + 1, start + 3 * 2 * n + 12345]
        try:
            for flag in (0, 1):
                kh.set_option("bsgs_base_check", flag)
                for key in keys:
                    pub = oracle.pubkey(key)
                    want = oracle.bsgs_search(b, pub, start, end, base_check=bool(flag))[0]
                    assert kh.bsgs_search(pub, start, end) == want, (flag, hex(key))
            kh.set_option("bsgs_base_check", 1)
            assert kh.bsgs_search(oracle.pubkey(start + 2 * n), start, end) == start + 2 * n
            kh.set_option("bsgs_base_check", 0)
            assert kh.bsgs_search(oracle.pubkey(start + 2 * n), start, end) == oracle.bsgs_search(b, oracle.pubkey(start + 2 * n), start, end)[0]
        finally:
            kh.set_option("bsgs_base_check", 0)
    finally:
        oracle.bsgs_free(b)


def test_k4096_tables_resident_in_one_gpu(kh, oracle):
    """bsgs -k 4096 (the reference's own BSGSD.md example: m = 2^34 baby points, 61.75 GB tier-1 bloom + the 64 GB prefix
    bitmap, everything resident in one B200's HBM): the reference's sizes, sampled baby points are members of tier 1, the
    puzzle-63 key of that example and keys planted in the first / last window of a 2^64 range are found, a miss is a miss."""
    kh.bsgs_build(1 << 44, 4096)
    try:
        d = kh.bsgs_describe()
        assert (d.m, d.m2, d.m3, d.aux) == (1 << 34, 1 << 29, 1 << 24, 1024)
        rnd = random.Random(4096)
        for j in [1, d.m] + [rnd.randrange(1, d.m + 1) for _ in range(4)]:
            xb = oracle.pubkey(j)[0].to_bytes(32, "big")
            bf, desc = kh.bsgs_export(1, xb[0]), d.tier[0]
            a = oracle.xxh64(xb, 0x59f2815b16f81798)
            b = oracle.xxh64(xb, a)
            assert all((bf[(((a + b * i) & (2**64 - 1)) % desc.bits) >> 3] >> ((((a + b * i) & (2**64 - 1)) % desc.bits) & 7)) & 1
                       for i in range(desc.hashes)), j
            del bf
        p63 = 0x7CCE5EFDACCF6808
        lo, hi = 1 << 64, 1 << 65
        cases = [(p63, 1 << 62, 1 << 63), (lo + 12345, lo, hi), (hi - 99, lo, hi), (hi + (1 << 50), lo, hi)]
        for key, a, b in cases:
            want = key if a <= key < b else None
            assert kh.bsgs_search(oracle.pubkey(key), a, b) == want, hex(key)
    finally:
        kh.bsgs_build(1 << 20, 1)          # give the 126 GB back to the rest of the session
