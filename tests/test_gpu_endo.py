"""GPU: endomorphism (-e, keyhunt.cpp:3408-3473 / :3556-3618) through the C ABI against the oracle, and through the
command-line driver against the records the unmodified reference binary wrote (tests/golden/scans_endo.json)."""
import json
import os
import random
import shutil
import subprocess
import tempfile

import pytest

import keyhunt_b200 as K
from _oracle import (BETA, BETA2, CRYPTO_BTC as O_BTC, CRYPTO_ETH as O_ETH, MODE_ADDRESS as O_ADDR, MODE_RMD160 as O_RMD,
                     MODE_XPOINT as O_XP, P_FIELD, SEARCH_BOTH, SEARCH_COMPRESS, SEARCH_UNCOMPRESS, be32)

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CLI = os.path.join(ROOT, "keyhunt_b200", "keyhunt-b200")
ENDO = {c["name"]: c for c in json.load(open(os.path.join(GOLD, "scans_endo.json")))}

CASES = [
    ("xpoint", K.MODE_XPOINT, K.CRYPTO_BTC, K.SEARCH_COMPRESS, O_XP, O_BTC, SEARCH_COMPRESS),
    ("comp", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_COMPRESS, O_RMD, O_BTC, SEARCH_COMPRESS),
    ("uncomp", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_UNCOMPRESS, O_RMD, O_BTC, SEARCH_UNCOMPRESS),
    ("both", K.MODE_ADDRESS, K.CRYPTO_BTC, K.SEARCH_BOTH, O_ADDR, O_BTC, SEARCH_BOTH),
    ("eth", K.MODE_ADDRESS, K.CRYPTO_ETH, K.SEARCH_COMPRESS, O_ADDR, O_ETH, SEARCH_COMPRESS),
]


@pytest.mark.parametrize("name,mode,crypto,search,omode,ocrypto,osearch", CASES)
def test_endomorphism_hits_equal_oracle(kh, oracle, name, mode, crypto, search, omode, ocrypto, osearch):
    rnd = random.Random(hash(name) & 0xFFFF)
    start, stride, n = 0x3000000000000321, 5, 1 << 14
    P = P_FIELD
    recs = []
    for j, i in enumerate(sorted({0, 512, 1023, 1024, n - 1} | {rnd.randrange(n) for _ in range(25)})):
        x, y = oracle.pubkey(start + i * stride)
        xv = [x, x * BETA % P, x * BETA2 % P][j % 3]
        yy = y if (j // 3) % 2 == 0 else P - y
        if name == "xpoint":
            recs.append(be32(xv)[:20])
        elif name == "eth":
            recs.append(oracle.eth_addr(xv, yy))
        elif name == "comp" or (name == "both" and j % 2):
            recs.append(oracle.hash160_comp(2 + (yy & 1), xv))
        else:
            recs.append(oracle.hash160_uncomp(xv, yy))
    recs += [rnd.randbytes(20) for _ in range(100)]
    kh.set_option("endomorphism", 1)
    try:
        kh.set_targets(mode, b"".join(recs), crypto=crypto, search=search)
        kh.scan(start, n, stride)
        got = kh.poll_hits()
    finally:
        kh.set_option("endomorphism", 0)
    t = oracle.targets_new(b"".join(recs))
    want = oracle.scan(t, omode, ocrypto, osearch, start, stride, n, endo=True)
    oracle.targets_free(t)
    assert sorted((h.index, h.kind, h.variant, h.key, h.matched) for h in got) == \
           sorted((h["index"], h["kind"], h["variant"], h["key"], h["matched"]) for h in want)
    assert len(got) >= 20
    for h in got:
        assert oracle.pubkey(h.key) == (h.pub_x, h.pub_y)


@pytest.mark.parametrize("name", sorted(ENDO))
def test_cli_endomorphism_records_equal_reference(name):
    """same command line as the recorded reference run (-e): identical KEYFOUNDKEYFOUND.txt records"""
    case = ENDO[name]
    d = tempfile.mkdtemp(prefix="khendo_")
    try:
        fn = os.path.join(d, "targets.txt")
        open(fn, "w").write("\n".join(case["targets"]) + "\n")
        args = case["args"].split()
        args[args.index("-f") + 1] = fn
        r = subprocess.run([CLI] + args + ["-q", "-t", "1"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r.returncode == 0 and "End" in r.stdout, r.stdout[-2000:]
        lines = open(os.path.join(d, "KEYFOUNDKEYFOUND.txt")).read().splitlines()
        per = 2 if name == "endo_eth" else 4
        assert sorted("|".join(lines[i:i + per]) for i in range(0, len(lines), per)) == case["records"]
    finally:
        shutil.rmtree(d, ignore_errors=True)
