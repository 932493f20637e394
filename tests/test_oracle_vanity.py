"""CPU: the oracle's -m vanity restatement (kho_b58tobin / kho_addvanity / vanityrmdmatch) against the reference:
tests/golden/vanity.json holds VANITYKEYFOUND.txt records written by the UNMODIFIED reference binary
(tests/golden/make_vanity_golden.py), and, where the harness is built, the reference's own b58tobin is called directly."""
import json
import os
import random

import pytest

from _oracle import CRYPTO_BTC, MODE_RMD160, SEARCH_BOTH, SEARCH_COMPRESS, SEARCH_UNCOMPRESS, Oracle, have_ref_harness

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vanity.json")))
SEARCH = {"compress": SEARCH_COMPRESS, "uncompress": SEARCH_UNCOMPRESS, "both": SEARCH_BOTH}


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: "+".join(c["prefixes"]) + "-" + c["search"])
def test_vanity_records_match_reference_binary(oracle, case):
    A, B, mn, counts = oracle.addvanity(case["prefixes"])
    assert (A.hex(), B.hex(), mn, counts) == (case["limits_a"], case["limits_b"], case["min_bytes"], case["per_prefix"])
    t = oracle.targets_new_vanity(A, B, mn)
    hits = oracle.scan(t, MODE_RMD160, CRYPTO_BTC, SEARCH[case["search"]], case["start"], 1, case["n_points"], max_hits=1 << 16)
    oracle.targets_free(t)
    assert sorted(["%x" % h["key"], h["kind"] != 2, h["matched"].hex()] for h in hits) == case["records"]


def test_addvanity_shapes(oracle):
    # a prefix that is a whole address has one zero-width interval; over-long and impossible prefixes add nothing
    assert oracle.addvanity(["1BgGZ9tcN4rm9KBzDn7KprQz87SZ26SAMH"])[3] == [0]       # 30 characters or more: refused (:6751)
    A, B, mn, c = oracle.addvanity(["1BgGZ9tcN4rm9KBzDn7KprQz87SZ2"])               # 29 characters: narrow intervals
    assert c[0] >= 1 and all(A[20 * i:20 * i + 20] <= B[20 * i:20 * i + 20] for i in range(c[0]))
    # two address lengths (33 and 34 characters) for an ordinary prefix, lower limit <= upper limit
    A, B, mn, c = oracle.addvanity(["1Bit"])
    assert c == [2] and all(A[20 * i:20 * i + 20] <= B[20 * i:20 * i + 20] for i in range(2))


@pytest.mark.ref
@pytest.mark.skipif(not have_ref_harness(), reason="oracle/_ref/libkh_ref.so not built")
def test_b58tobin_equals_reference_decoder(oracle):
    from _oracle import RefHarness
    r = RefHarness()
    rnd = random.Random(5)
    D = "123456789ABCDEFGHJKLMNPQRSTUVWXYZabcdefghijkmnopqrstuvwxyz"
    for _ in range(2000):
        t = "".join(rnd.choice(D) for _ in range(rnd.randrange(1, 36)))
        if rnd.random() < 0.4:
            t = "1" * rnd.randrange(1, 4) + t
        if rnd.random() < 0.05:
            t = t[:-1] + rnd.choice("0OIl+")
        for sz in (50, 25, 21, 4):
            a, b = oracle.b58tobin(t.encode(), sz), r.b58tobin(t.encode(), sz)
            assert a[0] == b[0] and (not a[0] or a == b), (t, sz)
