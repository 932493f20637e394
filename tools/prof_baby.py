#!/usr/bin/env python3
"""bsgs build at -k K (default 64: 2^28 baby steps) for ncu captures of kh_baby_kernel; prints the baby-step rate."""
import sys
sys.path.insert(0, ".")
import keyhunt_b200 as K
k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kh = K.KeyHunt(0)
kh.set_option("steps_per_launch", 1)
kh.bsgs_build(1 << 44, k)
s = kh.stats()
d = kh.bsgs_describe()
print("k", k, "m", d.m, "baby steps/s", d.m / s["walk_ms"] / 1e-3 / 1e9, "G  launches", s["walk_launches"])
