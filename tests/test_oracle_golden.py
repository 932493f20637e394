"""CPU: the oracle (oracle/kh_oracle.c) reproduces every golden vector the reference produced
(tests/golden/*.json, see tests/golden/README.md).  This is what pins the oracle."""
import hashlib
import json
import os
import random

import pytest

from _oracle import (CRYPTO_BTC, CRYPTO_ETH, MODE_ADDRESS, MODE_RMD160, MODE_XPOINT, SEARCH_BOTH, SEARCH_COMPRESS,
                     SEARCH_UNCOMPRESS)

GOLD = os.path.join(os.path.dirname(__file__), "golden")
PRIM = json.load(open(os.path.join(GOLD, "primitives.json")))
SCANS = {c["name"]: c for c in json.load(open(os.path.join(GOLD, "scans.json")))}
BSGS = json.load(open(os.path.join(GOLD, "bsgs.json")))


def I(s):
    return int(s, 16)


def test_field(oracle):
    for a, b, r in PRIM["fe_mul"]:
        assert oracle.fe_mul(I(a), I(b)) == I(r)
    for a, r in PRIM["fe_sqr"]:
        assert oracle.fe_sqr(I(a)) == I(r)
    for a, r in PRIM["fe_inv"]:
        assert oracle.fe_inv(I(a)) == I(r)


def test_pubkey_and_hashes(oracle):
    for k, x, y in PRIM["pubkey"]:
        assert oracle.pubkey(I(k)) == (I(x), I(y))
    for h in PRIM["hashes"]:
        x, y = oracle.pubkey(I(h["key"]))
        assert oracle.hash160_comp(2, x).hex() == h["c02"]
        assert oracle.hash160_comp(3, x).hex() == h["c03"]
        assert oracle.hash160_uncomp(x, y).hex() == h["unc"] == h["scalar_unc"]
        assert oracle.eth_addr(x, y).hex() == h["eth"]
        assert oracle.hash160_comp(2 + (y & 1), x).hex() == h["scalar_comp"]


def test_xxh64(oracle):
    for d, s, h20, h32 in PRIM["xxh64"]:
        d = bytes.fromhex(d)
        assert oracle.xxh64(d[:20], I(s)) == I(h20)
        assert oracle.xxh64(d, I(s)) == I(h32)


def test_bloom_sizing_and_images(oracle):
    for want in PRIM["bloom_sizing"]:
        h = oracle.bloom_new(want["entries"])
        assert oracle.bloom_desc(h) == want
        oracle.bloom_free(h)
    with pytest.raises(ValueError):
        oracle.bloom_new(999)          # bloom_init2 rejects entries < 1000 (bloom.cpp:157)
    for img in PRIM["bloom_images"]:
        rr = random.Random(img["seed"])
        h = oracle.bloom_new(img["entries"])
        items = [rr.randbytes(20) for _ in range(img["n20"])] + [rr.randbytes(32) for _ in range(img["n32"])]
        for it in items:
            oracle.bloom_add(h, it)
        assert hashlib.sha256(oracle.bloom_bytes(h)).hexdigest() == img["sha256"]
        assert all(oracle.bloom_check(h, it) for it in items)
        probes = [rr.randbytes(20 if img["n20"] else 32) for _ in range(3000)]
        assert [i for i, p in enumerate(probes) if oracle.bloom_check(h, p)] == img["false_positive_probe_indices"]
        oracle.bloom_free(h)


def test_batch_geometry(oracle):
    for b in PRIM["batches"]:
        raw = oracle.batch_points(I(b["base"]), I(b["stride"]), True)
        assert hashlib.sha256(raw).hexdigest() == b["sha256_xy"]
        assert raw[:32].hex() == b["x0"] and raw[512 * 64:512 * 64 + 32].hex() == b["x512"]


def _records(name):
    import keyhunt_b200 as K
    if name.startswith("planted"):
        return b"".join(bytes.fromhex(t) for t in SCANS[name]["targets"])
    fn, mode, crypto = {"rmd160_compress_1to32": ("1to32.rmd", K.MODE_RMD160, K.CRYPTO_BTC),
                        "address_compress_1to32": ("1to32.txt", K.MODE_ADDRESS, K.CRYPTO_BTC),
                        "address_eth_1to32": ("1to32.eth", K.MODE_ADDRESS, K.CRYPTO_ETH),
                        "xpoint_substracted40": ("substracted40.txt", K.MODE_XPOINT, K.CRYPTO_BTC)}[name]
    return K.parse_targets(open(os.path.join(GOLD, fn)), mode, crypto)


def test_target_parsers_agree():
    # base58 addresses and hash160 list of the same puzzles decode to the same records
    assert sorted(_records("address_compress_1to32")[i:i + 20] for i in range(0, 640, 20)) == \
           sorted(_records("rmd160_compress_1to32")[i:i + 20] for i in range(0, 640, 20))
    assert len(_records("address_eth_1to32")) == 32 * 20
    assert len(_records("xpoint_substracted40")) == 6003 * 20


@pytest.mark.parametrize("name,mode,crypto,search,start,n", [
    ("rmd160_compress_1to32", MODE_RMD160, CRYPTO_BTC, SEARCH_COMPRESS, 1, 1 << 24),
    ("address_eth_1to32", MODE_ADDRESS, CRYPTO_ETH, SEARCH_COMPRESS, 1, 1 << 24),
    ("planted_uncompress", MODE_RMD160, CRYPTO_BTC, SEARCH_UNCOMPRESS, 0x2000000000000000, 1 << 22),
    ("planted_both", MODE_RMD160, CRYPTO_BTC, SEARCH_BOTH, 0x2000000000000000, 1 << 22),
    ("planted_opposite_parity", MODE_RMD160, CRYPTO_BTC, SEARCH_COMPRESS, 0x2000000000000000, 1 << 22),
])
def test_scan_reproduces_reference_hits(oracle, name, mode, crypto, search, start, n):
    t = oracle.targets_new(_records(name))
    hits = oracle.scan(t, mode, crypto, search, start, 1, n, nthreads=8)
    oracle.targets_free(t)
    assert sorted(h["key"] for h in hits) == sorted(I(k) for k in SCANS[name]["keys"])
    assert SCANS["address_compress_1to32"]["keys"] == SCANS["rmd160_compress_1to32"]["keys"]
    if name == "planted_opposite_parity":   # the n-k rule (keyhunt.cpp:3629-3635)
        assert all(h["key"] > 2**255 for h in hits)


def test_xpoint_reproduces_reference_hits(oracle):
    """tests/substracted40.txt: the oracle finds the reference's two keys (and nothing else) in the 2^20-key
    chunks that contain them, and nothing in a few other chunks"""
    c = SCANS["xpoint_substracted40"]
    keys = [I(k) for k in c["keys"]]
    assert 0x800258a2ce in keys                      # README.md:416-437
    t = oracle.targets_new(_records("xpoint_substracted40"))
    chunks = sorted({k >> 20 << 20 for k in keys} | {0x8000000000, 0x8000400000, 0x800FF00000})
    got = []
    for s in chunks:
        got += [h["key"] for h in oracle.scan(t, MODE_XPOINT, CRYPTO_BTC, SEARCH_COMPRESS, s, 1, 1 << 20, nthreads=8)]
    oracle.targets_free(t)
    assert sorted(got) == sorted(keys)


def test_bsgs_build_and_search_reproduce_reference(oracle):
    b = oracle.bsgs_new(1 << 22, 2)
    try:
        p = oracle.bsgs_params(b)
        assert (p["m"], p["m2"], p["m3"], p["aux"]) == (4096, 128, 4, 1024)
        for tier, fn in [(1, "keyhunt_bsgs_4_4096.blm"), (2, "keyhunt_bsgs_6_128.blm"), (3, "keyhunt_bsgs_7_4.blm")]:
            f = BSGS["files"][fn]
            assert f["shard_checksums_ok"]
            d = oracle.bloom_desc(oracle.bsgs_bloom(b, tier, 0))
            assert (d["entries"], d["bits"], d["bytes"], d["hashes"]) == (f["entries"], f["bits"], f["shard_bytes"], f["hashes"])
            allb = b"".join(oracle.bloom_bytes(oracle.bsgs_bloom(b, tier, s)) for s in range(256))
            assert hashlib.sha256(allb).hexdigest() == f["sha256_all_shards"]
        raw = oracle.bsgs_table(b)
        ents = sorted((raw[i:i + 6].hex(), int.from_bytes(raw[i + 8:i + 16], "little")) for i in range(0, len(raw), 16))
        assert ents == [tuple(e) for e in BSGS["files"]["keyhunt_bsgs_2_4.tbl"]["entries"]]
        P = 2**256 - 2**32 - 977
        want = sorted(I(k) for k in BSGS["keys"])
        got = []
        for pk in BSGS["pubkeys"][:6]:
            x = I(pk[2:])
            y = pow((x**3 + 7) % P, (P + 1) // 4, P)
            if (y & 1) != (int(pk[:2], 16) & 1):
                y = P - y
            k, _, _ = oracle.bsgs_search(b, (x, y), 0x100000, 0x10000000000)
            got.append(k)
        assert sorted(got) == want[:6]
    finally:
        oracle.bsgs_free(b)
