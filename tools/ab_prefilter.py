import os, sys, time
sys.path.insert(0, ".")
import numpy as np
import keyhunt_b200 as K
kh = K.KeyHunt(0)
cases = [("xpoint", K.MODE_XPOINT, K.CRYPTO_BTC, K.SEARCH_COMPRESS), ("comp", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_COMPRESS),
         ("uncomp", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_UNCOMPRESS), ("both", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_BOTH),
         ("eth", K.MODE_ADDRESS, K.CRYPTO_ETH, K.SEARCH_COMPRESS)]
for ntg in [int(a) for a in sys.argv[1:]] or (32, 1024, 16384, 65536):
    recs = np.random.default_rng(1).integers(0, 256, size=ntg * 20, dtype=np.uint8).tobytes()
    for pre in (0, 1):
        kh.set_option("prefilter", pre)
        out = []
        for name, mode, crypto, search in cases:
            kh.set_targets(mode, recs, crypto=crypto, search=search)
            n = 1 << (33 if name == "xpoint" else 32)
            kh.scan(0x4000000000000000, 1 << 26)
            kh.stats(reset=True)
            kh.scan(0x4000000000000000, n)
            s = kh.stats(reset=True)
            out.append("%s %.0f" % (name, n / s["walk_ms"] / 1e3))
        print("targets", ntg, "prefilter", pre, " ".join(out), flush=True)
