#!/usr/bin/env python3
"""Small end-to-end pass over every kernel family (five scan kinds with planted keys, direct and binned / sliced BSGS builds with
their digests, a BSGS search), sized so that it also finishes under compute-sanitizer where that tool is available:
   [compute-sanitizer --tool memcheck] python tools/sanitize_small.py"""
import os, random, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import keyhunt_b200 as K
from _oracle import Oracle

o = Oracle()
rnd = random.Random(5)
kh = K.KeyHunt(0)
kh.set_option("threads_per_sm", 256)
start = 0x2000000000000000
n = 1 << 17
keys = [start + 5, start + n - 1]
for name, mode, crypto, search in (("both", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_BOTH), ("comp", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_COMPRESS),
                                   ("uncomp", K.MODE_RMD160, K.CRYPTO_BTC, K.SEARCH_UNCOMPRESS), ("eth", K.MODE_ADDRESS, K.CRYPTO_ETH, K.SEARCH_COMPRESS),
                                   ("xpoint", K.MODE_XPOINT, K.CRYPTO_BTC, K.SEARCH_COMPRESS)):
    recs = []
    for k in keys:
        x, y = o.pubkey(k)
        if name == "eth": recs.append(o.eth_addr(x, y))
        elif name == "xpoint": recs.append(x.to_bytes(32, "big")[:20])
        elif name == "uncomp": recs.append(o.hash160_uncomp(x, y))
        else: recs.append(o.hash160_comp(2 + (y & 1), x))
    recs += [rnd.randbytes(20) for _ in range(30)]
    kh.set_targets(mode, b"".join(recs), crypto=crypto, search=search)
    kh.scan(start, n)
    got = sorted(h.key for h in kh.poll_hits())
    assert got == keys, (name, got)
    print(name, "ok", flush=True)
for mode in (0, 2):
    kh.set_option("bsgs_binned_build", mode)
    if mode == 2:
        os.environ["KH_BABY_SLICE_KB"] = "4"
    kh.bsgs_build(1 << 22, 2)
    print("digests", [hex(kh.bsgs_digest(t)) for t in range(5)], flush=True)
    key = 0x8000000000 + 1234567
    assert kh.bsgs_search(o.pubkey(key), 0x8000000000, 0x8000000000 + (1 << 26)) == key
print("bsgs ok")
kh.close()
