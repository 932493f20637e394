#!/usr/bin/env python3
"""Quick throughput probe of every scan kind (device-timed walk only)."""
import os, random, sys, time
import numpy as np
sys.path.insert(0, ".")
import keyhunt_b200 as K

kh = K.KeyHunt(0)
print(kh.device_info())
rnd = random.Random(1)
ntg = int(os.environ.get("KH_NTG", "1024"))  # target-set size: 1024 = L1-resident bloom, 1e6 = L2, 5e7 = HBM
recs = np.random.default_rng(1).integers(0, 256, size=ntg * 20, dtype=np.uint8).tobytes()
print("targets", ntg)
cases = [("xpoint", K.MODE_XPOINT, K.CRYPTO_BTC, K.SEARCH_COMPRESS), ("comp", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_COMPRESS),
         ("uncomp", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_UNCOMPRESS), ("both", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_BOTH),
         ("eth", K.MODE_ADDRESS, K.CRYPTO_ETH, K.SEARCH_COMPRESS)]
tps = [int(a) for a in sys.argv[1:]] or [512]
for tp in tps:
    kh.set_option("threads_per_sm", tp)
    for name, mode, crypto, search in cases:
        kh.set_targets(mode, recs, crypto=crypto, search=search)
        n = 1 << (33 if name == "xpoint" else 32)
        kh.scan(0x4000000000000000, 1 << 26)  # warm
        kh.stats(reset=True)
        t0 = time.time()
        kh.scan(0x4000000000000000, n)
        wall = time.time() - t0
        s = kh.stats(reset=True)
        print(f"tp={tp} {name:7s} points={n} walk_ms={s['walk_ms']:.1f} setup_ms={s['setup_ms']:.2f} wall={wall*1e3:.1f}ms "
              f"Mpts/s={n/s['walk_ms']/1e3:.1f} launches={s['walk_launches']} T={s['walker_threads']}", flush=True)
