#!/usr/bin/env python3
"""Generates tests/golden/bsgsd_replies.json FROM the unmodified reference server (oracle/_ref/bsgsd, built by
oracle/Makefile from /root/reference): the request list of tests/test_gpu_bsgsd.py and the reference's replies.
Run in the build container:  python tests/golden/make_bsgsd_golden.py"""
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import test_gpu_bsgsd as T
from _oracle import Oracle

o = Oracle()
keys = {}
for k in (0x1234567, 0x1FFFFFF, 0x2000001, 0x2000000, 0x123456789, 1):
    x, y = o.pubkey(k)
    xb, yb = x.to_bytes(32, "big"), y.to_bytes(32, "big")
    keys[k] = ((b"03" if y & 1 else b"02") + xb.hex().encode(), b"04" + (xb + yb).hex().encode())
d = tempfile.mkdtemp(prefix="bsgsd_gold_")
srv = T.Server(T.REF_D, d, ["-6"])
out = []
try:
    for q in T.requests(keys):
        out.append({"request": q.decode("latin1"), "reply": T.mask(srv.ask(q)).decode("latin1")})
        assert srv.proc.poll() is None, ("reference server died on", q)
finally:
    srv.stop()
    shutil.rmtree(d, ignore_errors=True)
json.dump({"n": T.NK[1], "k": T.NK[3], "cases": out}, open(os.path.join(HERE, "bsgsd_replies.json"), "w"), indent=1)
for c in out:
    print(repr(c["request"][:40]), "->", repr(c["reply"][:60]))
