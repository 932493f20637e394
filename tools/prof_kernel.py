#!/usr/bin/env python3
"""One short scan of a given kind (for ncu section captures): python tools/prof_kernel.py comp|both|xpoint|uncomp|eth [log2 points] [targets]"""
import random, sys
sys.path.insert(0, ".")
import keyhunt_b200 as K
kind = sys.argv[1] if len(sys.argv) > 1 else "comp"
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 27
ntg = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
cases = {"xpoint": (K.MODE_XPOINT, K.CRYPTO_BTC, K.SEARCH_COMPRESS), "comp": (K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_COMPRESS),
         "uncomp": (K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_UNCOMPRESS), "both": (K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_BOTH),
         "eth": (K.MODE_ADDRESS, K.CRYPTO_ETH, K.SEARCH_COMPRESS)}
mode, crypto, search = cases[kind]
kh = K.KeyHunt(0)
kh.set_option("steps_per_launch", 1)
rnd = random.Random(1)
kh.set_targets(mode, b"".join(rnd.randbytes(20) for _ in range(ntg)), crypto=crypto, search=search)
for _ in range(3):
    kh.scan(0x4000000000000000, 1 << lg)
s = kh.stats()
print(kind, "Mpts/s", s["points"] / s["walk_ms"] / 1e3, "launches", s["walk_launches"])
