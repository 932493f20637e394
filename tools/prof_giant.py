#!/usr/bin/env python3
"""bsgs -k 512 build + one short giant-step search (for ncu application-replay captures of kh_giant_kernel)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import keyhunt_b200 as K
kh = K.KeyHunt(0)
kh.set_option("steps_per_launch", 2)
kh.bsgs_build(1 << 44, 512)
kh.stats(reset=True)
# a public key that is NOT in the range: plain sweep of 2^28 giant steps (2^15 windows)
G = (0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798, 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8)
r = kh.bsgs_search(G, 1 << 64, (1 << 64) + (1 << 15) * 2 * (1 << 44))
s = kh.stats()
print("found", r, "giant steps/s", s["points"] / s["walk_ms"] / 1e-3 / 1e9, "G  launches", s["walk_launches"], "tier1 pos", s["tier1_positives"])
