#!/usr/bin/env python3
"""Per-call cost of kh_bsgs_search at -k K (default 512) for calls of 2^lg giant steps: walk / set-up / refine device time and wall."""
import json, sys, time
sys.path.insert(0, ".")
import keyhunt_b200 as K
k = int(sys.argv[1]) if len(sys.argv) > 1 else 512
G = (0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798, 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8)
kh = K.KeyHunt(0)
t0 = time.time(); kh.bsgs_build(1 << 44, k); build_s = time.time() - t0
d = kh.bsgs_describe()
st = kh.stats(reset=True)
out = {"k": k, "build_wall_s": build_s, "build_walk_ms": st["walk_ms"], "build_sort_ms": st["aux_ms"], "calls": []}
win = 2 * (1 << 44)
for lg in (22, 24, 26, 28, 30, 32):
    nwin = (1 << lg) // d.aux
    start = (1 << 64) + lg * (1 << 60)
    kh.bsgs_search(G, start, start + nwin * win)        # warm
    kh.stats(reset=True)
    reps = 4
    t0 = time.time()
    for r in range(reps):
        a = start + (r + 1) * nwin * win
        kh.bsgs_search(G, a, a + nwin * win)
    wall = (time.time() - t0) / reps
    st = kh.stats(reset=True)
    out["calls"].append({"giant_steps_log2": lg, "wall_ms": wall * 1e3, "walk_ms": st["walk_ms"] / reps, "setup_ms": st["setup_ms"] / reps, "refine_ms": st["aux_ms"] / reps,
                         "tier1_positives": st["tier1_positives"] / reps, "launches": (st["walk_launches"] + st["other_launches"]) / reps,
                         "G_steps_per_s_wall": (st["points"] / reps) / wall / 1e9, "G_steps_per_s_kernel": st["points"] / (st["walk_ms"] * 1e-3) / 1e9, "T": st["walker_threads"]})
print(json.dumps(out, indent=1))
