"""GPU: -m vanity.  kh_set_vanity + kh_scan through the C ABI against the oracle's restatement of
addvanity / vanityrmdmatch (itself checked against the reference binary: tests/golden/vanity.json), and the CLI side by
side with the unmodified reference binary: identical VANITYKEYFOUND.txt records."""
import json
import os
import shutil
import subprocess
import tempfile

import pytest

import keyhunt_b200 as K
from _oracle import CRYPTO_BTC, MODE_RMD160, REF_BIN, SEARCH_BOTH, SEARCH_COMPRESS, SEARCH_UNCOMPRESS

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "keyhunt_b200", "keyhunt-b200")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "vanity.json")))


@pytest.mark.parametrize("search", [SEARCH_COMPRESS, SEARCH_UNCOMPRESS, SEARCH_BOTH])
@pytest.mark.parametrize("endo", [False, True])
def test_vanity_scan_equals_oracle(kh, oracle, search, endo):
    A, B, mn, counts = oracle.addvanity(["1Bit", "1Ab", "1zz", "12"])
    t = oracle.targets_new_vanity(A, B, mn)
    start, n = 0x6000000000000001, 1 << 17
    want = oracle.scan(t, MODE_RMD160, CRYPTO_BTC, search, start, 1, n, max_hits=1 << 18, endo=endo)
    oracle.targets_free(t)
    kh.set_option("endomorphism", int(endo))
    kh.set_option("hit_capacity", 1 << 18)        # "12" matches one address in ~22
    try:
        kh.set_vanity(A, B, search=search)
        kh.scan(start, n)
        got = kh.poll_hits()
    finally:
        kh.set_option("endomorphism", 0)
        kh.set_option("hit_capacity", 1 << 16)
    assert sorted((h.index, h.kind, h.variant, h.matched, h.key) for h in got) == \
        sorted((h["index"], h["kind"], h["variant"], h["matched"], h["key"]) for h in want)
    assert len(got) > 50


def test_vanity_golden_reference_records(kh, oracle):
    """tests/golden/vanity.json: VANITYKEYFOUND.txt records the unmodified reference binary wrote"""
    for case in GOLD["cases"]:
        A, B, mn, _ = oracle.addvanity(case["prefixes"])
        assert (A.hex(), B.hex(), mn) == (case["limits_a"], case["limits_b"], case["min_bytes"])
        search = {"compress": SEARCH_COMPRESS, "uncompress": SEARCH_UNCOMPRESS, "both": SEARCH_BOTH}[case["search"]]
        kh.set_vanity(A, B, search=search)
        kh.scan(case["start"], case["n_points"])
        got = sorted((h.key, h.kind != 2, h.matched.hex()) for h in kh.poll_hits())
        assert got == sorted((int(r[0], 16), r[1], r[2]) for r in case["records"]), case["prefixes"]


def test_vanity_after_targets_and_back(kh, oracle):
    """switching between a table search and a vanity search on one context"""
    A, B, mn, _ = oracle.addvanity(["1Ab"])
    x, y = oracle.pubkey(0x1234)
    rec = oracle.hash160_comp(2 + (y & 1), x)
    kh.set_targets(K.MODE_RMD160, rec, search=SEARCH_COMPRESS)
    kh.scan(0x1000, 4096)
    assert [h.key for h in kh.poll_hits()] == [0x1234]
    kh.set_vanity(A, B, search=SEARCH_COMPRESS)
    kh.scan(0x1000, 1 << 16)
    n_van = len(kh.poll_hits())
    assert n_van > 10
    kh.set_targets(K.MODE_RMD160, rec, search=SEARCH_COMPRESS)
    kh.scan(0x1000, 4096)
    assert [h.key for h in kh.poll_hits()] == [0x1234]
    with pytest.raises(K.KhError):
        kh.set_vanity(b"", b"", search=SEARCH_COMPRESS)


def test_vanity_overflow_is_reported(kh, oracle):
    """a dense prefix with a small hit buffer: the dropped hits are reported (KH_EOVERFLOW), never silently lost"""
    A, B, mn, _ = oracle.addvanity(["1"])
    kh.set_option("hit_capacity", 1024)
    try:
        kh.set_vanity(A, B, search=SEARCH_COMPRESS)
        kh.scan(1, 1 << 16)
        with pytest.raises(K.KhError) as e:
            kh.poll_hits()
        assert e.value.code == -5                     # KH_EOVERFLOW
        kh.poll_hits()
    finally:
        kh.set_option("hit_capacity", 1 << 16)


def _records(d):
    fn = os.path.join(d, "VANITYKEYFOUND.txt")
    if not os.path.exists(fn):
        return []
    L = open(fn).read().splitlines()
    return sorted("|".join(L[i:i + 4]) for i in range(0, len(L), 4))


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/keyhunt not built")
@pytest.mark.parametrize("args", [
    ["-v", "1Bit", "-v", "1Ab", "-l", "compress"],
    ["-v", "1Bit", "-v", "1Ab", "-l", "both"],
    ["-v", "1GoodBoy", "-v", "1zz", "-l", "uncompress"],
    ["-v", "1Bit", "-v", "1Ab", "-l", "both", "-e"],     # rare enough: every record costs the reference a file open + a scalar multiplication
    ["-f", "van.txt", "-l", "compress"],
])
def test_cli_vanity_records_identical_to_reference(args):
    g, r = tempfile.mkdtemp(prefix="khvan_gpu_"), tempfile.mkdtemp(prefix="khvan_ref_")
    try:
        for d in (g, r):
            open(os.path.join(d, "van.txt"), "w").write("1Bit\n1Ab\nnot*base58\n\n1BadBoy\n")
        common = ["-m", "vanity", "-r", "1:100000", "-n", "0x100000", "-q"] + args
        pg = subprocess.run([CLI] + common + ["-t", "1"], cwd=g, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        pr = subprocess.run([REF_BIN] + common + ["-s", "0", "-t", "1"], cwd=r, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        assert pg.returncode == pr.returncode == 0, (pg.stdout[-1500:], pr.stdout[-1500:])
        a, b = _records(g), _records(r)
        assert a == b and len(a) > 5, (len(a), len(b))
        for line in ("[+] Mode vanity",) + (() if "-f" in args else ("[+] Added Vanity search : " + args[1],)):
            assert line in pg.stdout and line in pr.stdout
    finally:
        shutil.rmtree(g, ignore_errors=True)
        shutil.rmtree(r, ignore_errors=True)
