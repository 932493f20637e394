"""GPU parity tests for the scan path: every call goes through the C ABI (libkh_b200.so) and is compared
bit-for-bit with the CPU oracle on the same seeded inputs."""
import os
import random

import pytest

import keyhunt_b200 as K
from _oracle import (CRYPTO_BTC as O_BTC, CRYPTO_ETH as O_ETH, MODE_ADDRESS as O_ADDR, MODE_RMD160 as O_RMD,
                     MODE_XPOINT as O_XP, N_ORDER, SEARCH_BOTH, SEARCH_COMPRESS, SEARCH_UNCOMPRESS, be32)

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_derive_matches_oracle(kh, oracle):
    rnd = random.Random(11)
    keys = [1, 2, 3, 7, 0xFFFFFFFF, 2**64 + 5, N_ORDER - 1, N_ORDER - 2] + [rnd.randrange(1, N_ORDER) for _ in range(300)]
    infos = kh.derive(keys)
    for k, info in zip(keys, infos):
        x, y = oracle.pubkey(k)
        assert (info.pub_x, info.pub_y) == (x, y), hex(k)
        assert info.h160_comp == oracle.hash160_comp(2 + (y & 1), x)
        assert info.h160_uncomp == oracle.hash160_uncomp(x, y)
        assert info.eth == oracle.eth_addr(x, y)


def _hits_set(hits):
    return sorted((h.index, h.kind, h.key, h.matched) for h in hits)


def _ohits_set(hits):
    return sorted((h["index"], h["kind"], h["key"], h["matched"]) for h in hits)


def _planted(oracle, rnd, start, stride, n_points, kind, count):
    """targets planted inside the range for every hit kind + random decoys"""
    recs, idxs = [], sorted(set([0, 1, 511, 512, 513, 1023, 1024, n_points - 1] + [rnd.randrange(n_points) for _ in range(count)]))
    for j, i in enumerate(idxs):
        x, y = oracle.pubkey((start + i * stride) % N_ORDER)
        if kind == "xpoint":
            recs.append(be32(x)[:20])
        elif kind == "eth":
            recs.append(oracle.eth_addr(x, y))
        elif kind == "comp":
            # mix: real prefix, and the opposite prefix (a target whose key is n-k)
            pre = 2 + (y & 1)
            recs.append(oracle.hash160_comp(pre if j % 3 else 5 - pre, x))
        elif kind == "uncomp":
            recs.append(oracle.hash160_uncomp(x, y))
        else:  # both
            recs.append(oracle.hash160_uncomp(x, y) if j % 2 else oracle.hash160_comp(2 + (y & 1), x))
    recs += [rnd.randbytes(20) for _ in range(200)]
    rnd.shuffle(recs)
    return b"".join(recs)


CASES = [
    ("xpoint", K.MODE_XPOINT, K.CRYPTO_BTC, K.SEARCH_COMPRESS, O_XP, O_BTC, SEARCH_COMPRESS),
    ("comp", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_COMPRESS, O_RMD, O_BTC, SEARCH_COMPRESS),
    ("uncomp", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_UNCOMPRESS, O_RMD, O_BTC, SEARCH_UNCOMPRESS),
    ("both", K.MODE_ADDRESS, K.CRYPTO_BTC, K.SEARCH_BOTH, O_ADDR, O_BTC, SEARCH_BOTH),
    ("eth", K.MODE_ADDRESS, K.CRYPTO_ETH, K.SEARCH_COMPRESS, O_ADDR, O_ETH, SEARCH_COMPRESS),
]


@pytest.mark.parametrize("name,mode,crypto,search,omode,ocrypto,osearch", CASES)
@pytest.mark.parametrize("start,stride,n_points", [(1, 1, 1 << 16), (0x2000000000000000, 1, 3 * 1024),
                                                   (0xDEADBEEF12345, 977, 1 << 14)])
def test_scan_matches_oracle(kh, oracle, name, mode, crypto, search, omode, ocrypto, osearch, start, stride, n_points):
    rnd = random.Random(hash((name, start, stride)) & 0xFFFF)
    recs = _planted(oracle, rnd, start, stride, n_points, name, 12)
    kh.set_targets(mode, recs, crypto=crypto, search=search)
    # device-built bloom image and sorted table equal the oracle's (bloom_add / _sort parity)
    t = oracle.targets_new(recs)
    desc, bits = kh.get_bloom()
    assert desc.as_dict() == oracle.bloom_desc(oracle.targets_bloom(t))
    assert bits == oracle.bloom_bytes(oracle.targets_bloom(t))
    assert kh.get_table() == oracle.targets_table(t)
    kh.scan(start, n_points, stride)
    got = kh.poll_hits()
    want = oracle.scan(t, omode, ocrypto, osearch, start, stride, n_points)
    oracle.targets_free(t)
    assert len(want) >= 8
    assert _hits_set(got) == _ohits_set(want)
    for h in got:  # reported key really owns the reported public key
        assert oracle.pubkey(h.key) == (h.pub_x, h.pub_y)


def test_golden_puzzles_1_to_32_first_2pow24(kh):
    """reference fixture tests/1to32.rmd; hit list pinned by the unmodified reference binary
    (`keyhunt -m rmd160 -f tests/1to32.rmd -r 1:FFFFFF -l compress`, see tests/golden/README.md)"""
    recs = K.parse_targets(open(os.path.join(GOLD, "1to32.rmd")), K.MODE_RMD160)
    assert len(recs) == 32 * 20
    kh.set_targets(K.MODE_RMD160, recs, search=K.SEARCH_COMPRESS)
    kh.scan(1, 1 << 24)
    keys = sorted(h.key for h in kh.poll_hits())
    assert keys == [0x1, 0x3, 0x7, 0x8, 0x15, 0x31, 0x4c, 0xe0, 0x1d3, 0x202, 0x483, 0xa7b, 0x1460, 0x2930, 0x68f3,
                    0xc936, 0x1764f, 0x3080d, 0x5749f, 0xd2c55, 0x1ba534, 0x2de40f, 0x556e52, 0xdc2a04]


def test_multi_launch_and_tail(kh, oracle):
    """range that needs several launches per thread and ends mid-wave"""
    kh.set_option("steps_per_launch", 2)
    kh.set_option("threads_per_sm", 64)
    try:
        rnd = random.Random(5)
        start, n_points = 0x7000000000000123, 1024 * 40000 + 1024 * 7
        idxs = [0, n_points - 1, n_points // 2] + [rnd.randrange(n_points) for _ in range(20)]
        recs = b"".join(be32(oracle.pubkey(start + i)[0])[:20] for i in idxs)
        kh.set_targets(K.MODE_XPOINT, recs)
        kh.scan(start, n_points)
        got = sorted(h.index for h in kh.poll_hits())
        assert got == sorted(set(idxs))
    finally:
        kh.set_option("steps_per_launch", 16)
        kh.set_option("threads_per_sm", 4096)


def test_empty_and_duplicate_targets(kh, oracle):
    # empty target set: bloom of 10,000 entries with no bit set -> no hits, no errors (N = 0 after an all-invalid file)
    kh.set_targets(K.MODE_XPOINT, b"")
    d, bits = kh.get_bloom()
    assert d.entries == 10000 and not any(bits)
    kh.scan(12345, 4096)
    assert kh.poll_hits() == []
    # duplicate records: the table keeps both, a key is still reported once per matching hash
    x, _ = oracle.pubkey(0x5000 + 7)
    rec = be32(x)[:20]
    kh.set_targets(K.MODE_XPOINT, rec * 3 + bytes(20))
    assert kh.get_table() == bytes(20) + rec * 3
    kh.scan(0x5000, 1024)
    hits = kh.poll_hits()
    assert [(h.index, h.key) for h in hits] == [(7, 0x5007)]


def test_sorted_and_unsorted_target_sets_give_the_same_tables(kh, oracle):
    """kh_set_targets sends an already sorted record set to the device as it is (the reference's addressTable arrives sorted) and sorts
    any other on the host first; either way the words are packed on the device (kh_table_pack): same table, bloom, hits"""
    rnd = random.Random(77)
    start = 0x7000000
    keys = [start + 3, start + 2047]
    recs = [be32(oracle.pubkey(k)[0])[:20] for k in keys] + [rnd.randbytes(20) for _ in range(3000)]
    recs += [bytes(20), b"\xff" * 20, recs[5]]                      # extremes and a duplicate
    seen = []
    for order in (sorted(recs), recs, sorted(recs, reverse=True)):
        kh.set_targets(K.MODE_XPOINT, b"".join(order))
        d, bits = kh.get_bloom()
        kh.scan(start, 4096)
        seen.append((kh.get_table(), d.as_dict(), bits, [(h.index, h.key) for h in kh.poll_hits()]))
    assert seen[0][0] == b"".join(sorted(recs))
    assert seen[0] == seen[1] == seen[2]
    assert seen[0][3] == [(3, keys[0]), (2047, keys[1])]


def test_hit_buffer_overflow_is_reported(kh, oracle):
    """every point of the range is a target: more hits than the device hit buffer holds -> KH_EOVERFLOW (loud, not silent)"""
    start, n = 0x7000, 2048
    raw = b"".join(oracle.batch_points(start + b * 1024, 1, False) for b in range(n // 1024))
    recs = b"".join(raw[64 * i:64 * i + 20] for i in range(n))
    kh.set_targets(K.MODE_XPOINT, recs)
    kh.scan(start, n)
    assert sorted(h.index for h in kh.poll_hits()) == list(range(n))      # default capacity holds them all
    kh.set_option("hit_capacity", 16)
    try:
        kh.scan(start, n)
        with pytest.raises(K.KhError) as ei:
            kh.poll_hits()
        assert ei.value.code == -5
    finally:
        kh.set_option("hit_capacity", 1 << 16)
        kh.poll_hits()


def test_bad_arguments(kh):
    kh.set_targets(K.MODE_XPOINT, bytes(20))
    for bad in (0, 1000, 1025):
        with pytest.raises(K.KhError) as ei:
            kh.scan(1, bad)
        assert ei.value.code == -2
    with pytest.raises(K.KhError):
        kh.scan(1, 1024, stride=0)
    with pytest.raises(K.KhError):
        kh.bsgs_build(1 << 21, 1)      # not an even power of two
    with pytest.raises(K.KhError):
        kh.set_targets(K.MODE_BSGS, bytes(20))


def test_c3_full_size_planted_only(kh):
    """BASELINE config 3 at full size: 2^36 keys against 10^6 x-coordinates (999,968 random + 32 planted, bloom 3.6 MB,
    table 20 MB).  Size-independent property: exactly the planted keys are reported, nothing else."""
    rnd = random.Random(3)
    start, n = 0x4000000000000000, 1 << 36
    idx = sorted({0, n - 1} | {rnd.randrange(n) for _ in range(30)})
    infos = kh.derive([start + i for i in idx])
    planted = [be32(info.pub_x)[:20] for info in infos]
    recs = planted + [rnd.randbytes(20) for _ in range(1000000 - len(planted))]
    rnd.shuffle(recs)
    kh.set_targets(K.MODE_XPOINT, b"".join(recs))
    d, _ = kh.get_bloom()
    assert (d.entries, d.bits, d.bytes, d.hashes) == (1000000, 28755175, 3594397, 20)
    got = []
    for s in range(16):
        kh.scan(start + s * (n // 16), n // 16)
        got += [h.key for h in kh.poll_hits()]
    assert sorted(got) == [start + i for i in idx]


def test_c2_quarter_size_planted_only(kh):
    """BASELINE config 2 (rmd160 -l both, 1,024 targets) on 2^34 keys: compressed and uncompressed planted keys incl. the
    first and last key of the range and both Y parities; exactly those are reported."""
    rnd = random.Random(2)
    start, n = 0x2000000000000000, 1 << 34
    idx = sorted({0, n - 1} | {rnd.randrange(n) for _ in range(22)})
    infos = kh.derive([start + i for i in idx])
    planted = [(info.h160_uncomp if j % 2 else info.h160_comp) for j, info in enumerate(infos)]
    assert {info.pub_y & 1 for info in infos} == {0, 1}
    recs = planted + [rnd.randbytes(20) for _ in range(1000)]
    kh.set_targets(K.MODE_RMD160, b"".join(recs), search=K.SEARCH_BOTH)
    kh.scan(start, n)
    hits = kh.poll_hits()
    assert sorted(h.key for h in hits) == [start + i for i in idx]
    assert {h.kind for h in hits} == {K.HIT_COMP02, K.HIT_COMP03, K.HIT_UNCOMP}


@pytest.mark.parametrize("n_decoys", [0, 1000, 300000])
def test_prefilter_never_changes_the_hits(kh, oracle, n_decoys):
    """the exact prefix bitmap in front of the bloom (option "prefilter", on by default; bitmap 2^16 .. 2^27 bits here): same hits
    with it, without it (all-ones bitmap), and from the oracle, for every scan kind incl. targets that share their first bytes"""
    import numpy as np
    rnd = random.Random(77 + n_decoys)
    start, stride, n_points = 0x7000000000000123, 1, 1 << 15
    decoys = np.random.default_rng(n_decoys).integers(0, 256, size=n_decoys * 20, dtype=np.uint8).tobytes()
    for name, mode, crypto, search, omode, ocrypto in (("both", K.MODE_RMD160, K.CRYPTO_BTC, SEARCH_BOTH, O_RMD, O_BTC),
                                                       ("xpoint", K.MODE_XPOINT, K.CRYPTO_BTC, SEARCH_COMPRESS, O_XP, O_BTC),
                                                       ("eth", K.MODE_ADDRESS, K.CRYPTO_ETH, SEARCH_COMPRESS, O_ADDR, O_ETH)):
        recs = []
        for i in sorted({0, 1, 1023, 1024, n_points - 1} | {rnd.randrange(n_points) for _ in range(40)}):
            x, y = oracle.pubkey(start + i * stride)
            if name == "xpoint":
                recs.append(be32(x)[:20])
            elif name == "eth":
                recs.append(oracle.eth_addr(x, y))
            else:
                recs.append(oracle.hash160_uncomp(x, y) if i % 2 else oracle.hash160_comp(2 + (y & 1), x))
        near = [r[:3] + rnd.randbytes(17) for r in recs[:10]]          # same first 24 bits as a real target, not targets
        blob = b"".join(recs + near) + decoys
        t = oracle.targets_new(blob)
        want = _ohits_set(oracle.scan(t, omode, ocrypto, search, start, stride, n_points))
        oracle.targets_free(t)
        got = {}
        try:
            for pf in (1, 0):
                kh.set_option("prefilter", pf)
                kh.set_targets(mode, blob, crypto=crypto, search=search)
                kh.scan(start, n_points, stride)
                got[pf] = _hits_set(kh.poll_hits())
        finally:
            kh.set_option("prefilter", 1)
        assert got[1] == got[0] == want, name
        assert len(want) >= 40
