"""GPU vs the golden vectors the reference produced (tests/golden/*.json): primitives through kh_derive,
full hit lists of the reference binary on its own fixtures, BSGS -S files and found keys."""
import hashlib
import json
import os

import pytest

import keyhunt_b200 as K

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
PRIM = json.load(open(os.path.join(GOLD, "primitives.json")))
SCANS = {c["name"]: c for c in json.load(open(os.path.join(GOLD, "scans.json")))}
BSGS = json.load(open(os.path.join(GOLD, "bsgs.json")))


def I(s):
    return int(s, 16)


def test_derive_matches_reference_vectors(kh):
    keys = [I(h["key"]) for h in PRIM["hashes"]]
    infos = kh.derive(keys)
    pub = {I(k): (I(x), I(y)) for k, x, y in PRIM["pubkey"]}
    for h, info in zip(PRIM["hashes"], infos):
        assert (info.pub_x, info.pub_y) == pub[I(h["key"])]
        assert info.h160_comp.hex() == h["scalar_comp"]
        assert info.h160_uncomp.hex() == h["unc"]
        assert info.eth.hex() == h["eth"]


def test_bloom_sizing_matches_reference():
    for want in PRIM["bloom_sizing"]:
        assert K.bloom_params(want["entries"]).as_dict() == want


CASES = [
    ("rmd160_compress_1to32", "1to32.rmd", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_COMPRESS, 1, 1 << 24),
    ("address_compress_1to32", "1to32.txt", K.MODE_ADDRESS, K.CRYPTO_BTC, K.SEARCH_COMPRESS, 1, 1 << 24),
    ("address_eth_1to32", "1to32.eth", K.MODE_ADDRESS, K.CRYPTO_ETH, K.SEARCH_COMPRESS, 1, 1 << 24),
    ("xpoint_substracted40", "substracted40.txt", K.MODE_XPOINT, K.CRYPTO_BTC, K.SEARCH_COMPRESS, 0x8000000000, 1 << 28),
    ("planted_uncompress", None, K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_UNCOMPRESS, 0x2000000000000000, 1 << 22),
    ("planted_both", None, K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_BOTH, 0x2000000000000000, 1 << 22),
    ("planted_opposite_parity", None, K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_COMPRESS, 0x2000000000000000, 1 << 22),
]


@pytest.mark.parametrize("name,fn,mode,crypto,search,start,n", CASES)
def test_scan_equals_reference_binary(kh, name, fn, mode, crypto, search, start, n):
    """same flags as the reference run recorded in scans.json -> the same private keys"""
    if fn:
        recs = K.parse_targets(open(os.path.join(GOLD, fn)), mode, crypto)
    else:
        recs = b"".join(bytes.fromhex(t) for t in SCANS[name]["targets"])
    kh.set_targets(mode, recs, crypto=crypto, search=search)
    kh.scan(start, n)
    assert sorted(h.key for h in kh.poll_hits()) == sorted(I(k) for k in SCANS[name]["keys"])


def test_c1_full_sweep_finds_exactly_the_32_puzzle_keys(kh):
    """BASELINE config 1: -m address -f tests/1to32.txt -r 1:FFFFFFFF -l compress scans keys 1..2^32 (chunk
    overshoot) and reports exactly the 32 keys listed in SURVEY §8c"""
    recs = K.parse_targets(open(os.path.join(GOLD, "1to32.txt")), K.MODE_ADDRESS)
    kh.set_targets(K.MODE_ADDRESS, recs, search=K.SEARCH_COMPRESS)
    kh.scan(1, 1 << 32)
    want = ("1 3 7 8 15 31 4c e0 1d3 202 483 a7b 1460 2930 68f3 c936 1764f 3080d 5749f d2c55 1ba534 2de40f 340326e 556e52 "
            "6ac3875 dc2a04 1fa5ee5 d916ce8 17e2551e 3d94cd64 7d4fe747 b862a62e").split()
    assert sorted(h.key for h in kh.poll_hits()) == sorted(I(k) for k in want)


def test_bsgs_equals_reference_files_and_keys(kh):
    kh.bsgs_build(1 << 22, 2)
    for tier, fn in [(1, "keyhunt_bsgs_4_4096.blm"), (2, "keyhunt_bsgs_6_128.blm"), (3, "keyhunt_bsgs_7_4.blm")]:
        f = BSGS["files"][fn]
        allb = b"".join(kh.bsgs_export(tier, s) for s in range(256))
        assert hashlib.sha256(allb).hexdigest() == f["sha256_all_shards"]
    raw = kh.bsgs_export(0)
    ents = sorted((raw[i:i + 6].hex(), int.from_bytes(raw[i + 8:i + 16], "little")) for i in range(0, len(raw), 16))
    assert ents == [tuple(e) for e in BSGS["files"]["keyhunt_bsgs_2_4.tbl"]["entries"]]
    P = 2**256 - 2**32 - 977
    got = []
    for pk in BSGS["pubkeys"]:
        x = I(pk[2:])
        y = pow((x**3 + 7) % P, (P + 1) // 4, P)
        if (y & 1) != (int(pk[:2], 16) & 1):
            y = P - y
        got.append(kh.bsgs_search((x, y), 0x100000, 0x10000000000))
    assert sorted(got) == sorted(I(k) for k in BSGS["keys"])
