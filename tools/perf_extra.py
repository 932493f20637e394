#!/usr/bin/env python3
"""Throughput of the scan variants outside the headline configs: -m vanity and -e (device-timed walk only)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import keyhunt_b200 as K
from _oracle import Oracle
o = Oracle()
kh = K.KeyHunt(0)
A, B, mn, _ = o.addvanity(["1Bitcoin", "1Satoshi"])
recs = np.random.default_rng(1).integers(0, 256, size=1024 * 20, dtype=np.uint8).tobytes()
def run(label, n=1 << 32):
    kh.scan(0x4000000000000000, 1 << 26); kh.poll_hits(); kh.stats(reset=True)
    kh.scan(0x4000000000000000, n); kh.poll_hits()
    s = kh.stats(reset=True)
    print("%-28s %8.1f Mpoints/s" % (label, n / s["walk_ms"] / 1e3), flush=True)
for name, search in (("compress", K.SEARCH_COMPRESS), ("uncompress", K.SEARCH_UNCOMPRESS), ("both", K.SEARCH_BOTH)):
    kh.set_vanity(A, B, search=search); run("vanity -l " + name)
kh.set_option("endomorphism", 1)
for name, mode, crypto, search in (("rmd160 -l both -e", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_BOTH), ("rmd160 -l compress -e", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_COMPRESS),
                                   ("address eth -e", K.MODE_ADDRESS, K.CRYPTO_ETH, K.SEARCH_COMPRESS), ("xpoint -e", K.MODE_XPOINT, K.CRYPTO_BTC, K.SEARCH_COMPRESS)):
    kh.set_targets(mode, recs, crypto=crypto, search=search); run(name, 1 << 31)
kh.set_vanity(A, B, search=K.SEARCH_COMPRESS); run("vanity -l compress -e", 1 << 31)
