#!/usr/bin/env python3
"""Generates tests/golden/vanity.json FROM the unmodified reference binary (oracle/_ref/keyhunt): for each case the
VANITYKEYFOUND.txt records of `keyhunt -m vanity -v ... -r start:end`, plus the interval limits the oracle's addvanity
restatement derives for the same prefixes (checked here: the oracle scan over those limits reproduces the reference's
records exactly, otherwise the script fails).  Run in the build container: python tests/golden/make_vanity_golden.py"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _oracle import CRYPTO_BTC, MODE_RMD160, REF_BIN, SEARCH_BOTH, SEARCH_COMPRESS, SEARCH_UNCOMPRESS, Oracle

o = Oracle()
CASES = [(["1Bit", "1Ab"], "compress", 1, 0x100000), (["1Bit", "1zz"], "both", 0x100001, 0x100000), (["1GoodBoy", "1Go"], "uncompress", 1, 0x100000),
         (["3P9a", "11111"], "both", 1, 0x100000)]
out = []
for prefixes, search, start, n in CASES:
    d = tempfile.mkdtemp(prefix="vangold_")
    try:
        cmd = [REF_BIN, "-m", "vanity", "-l", search, "-r", "%x:%x" % (start, start + n - 1), "-n", "0x100000", "-q", "-s", "0", "-t", "1"]
        for p in prefixes:
            cmd += ["-v", p]
        subprocess.run(cmd, cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True, timeout=900)
        L = open(os.path.join(d, "VANITYKEYFOUND.txt")).read().splitlines()
    finally:
        shutil.rmtree(d, ignore_errors=True)
    recs = sorted([L[i].split()[-1], len(L[i + 1].split()[-1]) == 66, L[i + 3].split()[-1]] for i in range(0, len(L), 4))
    A, B, mn, counts = o.addvanity(prefixes)
    t = o.targets_new_vanity(A, B, mn)
    s = {"compress": SEARCH_COMPRESS, "uncompress": SEARCH_UNCOMPRESS, "both": SEARCH_BOTH}[search]
    hits = o.scan(t, MODE_RMD160, CRYPTO_BTC, s, start, 1, n, max_hits=1 << 20)
    o.targets_free(t)
    mine = sorted(["%x" % h["key"], h["kind"] != 2, h["matched"].hex()] for h in hits)
    assert mine == recs, (prefixes, len(mine), len(recs))
    out.append(dict(prefixes=prefixes, search=search, start=start, n_points=n, limits_a=A.hex(), limits_b=B.hex(), min_bytes=mn,
                    per_prefix=counts, records=recs))
    print(prefixes, search, len(recs), "records, oracle == reference")
json.dump({"cases": out}, open(os.path.join(HERE, "vanity.json"), "w"))
