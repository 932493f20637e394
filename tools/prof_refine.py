#!/usr/bin/env python3
"""bsgs -k 512 build + search of a PLANTED key (window 37): kh_refine_kernel runs on its tier-1 positive (ncu capture)."""
import random, sys
sys.path.insert(0, ".")
import keyhunt_b200 as K
kh = K.KeyHunt(0)
kh.bsgs_build(1 << 44, 512)
kh.stats(reset=True)
key = (1 << 64) + 37 * (1 << 45) + random.Random(4).randrange(1 << 45)
pub = kh.derive([key])[0]
r = kh.bsgs_search((pub.pub_x, pub.pub_y), 1 << 64, 1 << 65)
s = kh.stats()
print("found", r == key, "giant steps", s["points"], "tier1 positives", s["tier1_positives"], "aux_ms", s["aux_ms"])
