"""CPU: the oracle's -e (endomorphism) restatement reproduces, record for record, what the unmodified reference binary
wrote to KEYFOUNDKEYFOUND.txt on planted endomorphic targets (tests/golden/scans_endo.json) — including the reference's
ETH slot-4 quirk (keyhunt.cpp:3534)."""
import hashlib
import json
import os

import pytest

from _oracle import (CRYPTO_BTC, CRYPTO_ETH, HIT_COMP02, HIT_COMP03, HIT_ETH, MODE_ADDRESS, MODE_RMD160, MODE_XPOINT,
                     SEARCH_BOTH, SEARCH_COMPRESS, SEARCH_UNCOMPRESS, be32)

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ENDO = {c["name"]: c for c in json.load(open(os.path.join(GOLD, "scans_endo.json")))}
B58 = "123456789ABCDEFGHJKLMNPQRSTUVWXYZabcdefghijkmnopqrstuvwxyz"


def b58check_address(h160):
    d = b"\0" + h160
    d += hashlib.sha256(hashlib.sha256(d).digest()).digest()[:4]
    v = int.from_bytes(d, "big")
    s = ""
    while v:
        v, r = divmod(v, 58)
        s = B58[r] + s
    return "1" * (len(d) - len(d.lstrip(b"\0"))) + s


def record(oracle, hit):
    """what writekey / writekeyeth (keyhunt.cpp:6891, :6925) print for this hit"""
    key = hit["key"]
    x, y = oracle.pubkey(key)
    if hit["kind"] == HIT_ETH:
        return "Private Key: %x|address: 0x%s" % (key, oracle.eth_addr(x, y).hex())
    if hit["kind"] in (HIT_COMP02, HIT_COMP03):
        pub, h = "%02x%064x" % (2 + (y & 1), x), oracle.hash160_comp(2 + (y & 1), x)
    else:
        pub, h = "04%064x%064x" % (x, y), oracle.hash160_uncomp(x, y)
    return "Private Key: %x|pubkey: %s|Address %s|rmd160 %s" % (key, pub, b58check_address(h), h.hex())


def targets_of(case):
    out = []
    for t in case["targets"]:
        t = t[2:] if t.startswith("0x") else t
        out.append(bytes.fromhex(t)[:20])
    return b"".join(out)


CASES = [("endo_compress", MODE_RMD160, CRYPTO_BTC, SEARCH_COMPRESS), ("endo_uncompress", MODE_RMD160, CRYPTO_BTC, SEARCH_UNCOMPRESS),
         ("endo_both", MODE_RMD160, CRYPTO_BTC, SEARCH_BOTH), ("endo_eth", MODE_ADDRESS, CRYPTO_ETH, SEARCH_COMPRESS),
         ("endo_xpoint", MODE_XPOINT, CRYPTO_BTC, SEARCH_COMPRESS)]


@pytest.mark.parametrize("name,mode,crypto,search", CASES)
def test_oracle_endomorphism_records_equal_reference(oracle, name, mode, crypto, search):
    case = ENDO[name]
    t = oracle.targets_new(targets_of(case))
    hits = oracle.scan(t, mode, crypto, search, 0x3000000000000000, 1, 1 << 21, nthreads=8, endo=True)
    oracle.targets_free(t)
    assert sorted(record(oracle, h) for h in hits) == case["records"]
    assert len(hits) == 24
