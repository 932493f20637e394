#!/usr/bin/env python3
"""C4 at full size: bsgs -k 512 (n = 2^44): build time, size-independent checks, time-to-find, sweep rate."""
import json, random, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import keyhunt_b200 as K
from _oracle import Oracle

k = int(sys.argv[1]) if len(sys.argv) > 1 else 512
o = Oracle()
kh = K.KeyHunt(0)
import os
if os.environ.get('KH_BSGS_PREFILTER'):
    kh.set_option('bsgs_prefilter', int(os.environ['KH_BSGS_PREFILTER']))
if os.environ.get('KH_TPS'):
    kh.set_option('threads_per_sm', int(os.environ['KH_TPS']))
t0 = time.time()
kh.bsgs_build(1 << 44, k)
build_wall = time.time() - t0
st = kh.stats(reset=True)
d = kh.bsgs_describe()
res = {"k": k, "m": d.m, "m2": d.m2, "m3": d.m3, "aux": d.aux, "tier1_bytes_total": d.tier[0].bytes * 256,
       "build_wall_s": build_wall, "build_walk_ms": st["walk_ms"], "build_aux_ms": st["aux_ms"], "baby_steps_per_s": d.m / (st["walk_ms"] * 1e-3)}
# membership: random baby points must be present in all tiers they belong to
rnd = random.Random(4)
shards = {}
def check(tier, x):
    xb = x.to_bytes(32, "big")
    key = (tier, xb[0])
    if key not in shards:
        shards[key] = kh.bsgs_export(tier, xb[0])
    bf, desc = shards[key], d.tier[tier - 1]
    a = o.xxh64(xb, 0x59f2815b16f81798); b = o.xxh64(xb, a)
    for i in range(desc.hashes):
        bit = ((a + b * i) & (2**64 - 1)) % desc.bits
        if not (bf[bit >> 3] >> (bit & 7)) & 1:
            return False
    return True
ok1 = all(check(1, o.pubkey(j)[0]) for j in [1, d.m] + [rnd.randrange(1, d.m + 1) for _ in range(12)])
shards.clear()
ok2 = all(check(2, o.pubkey(j)[0]) for j in [1, d.m2] + [rnd.randrange(1, d.m2 + 1) for _ in range(12)])
ok3 = all(check(3, o.pubkey(j)[0]) for j in [1, d.m3] + [rnd.randrange(1, d.m3 + 1) for _ in range(12)])
neg = sum(check(1, o.pubkey(d.m + 1 + rnd.randrange(1 << 40))[0]) for _ in range(12))
shards.clear()
tab = kh.bsgs_export(0)
ents = [(tab[i:i + 6], int.from_bytes(tab[i + 8:i + 16], "little")) for i in range(0, len(tab), 16)]
res.update(members_tier1=ok1, members_tier2=ok2, members_tier3=ok3, nonmember_positives_of_12=neg,
           table_sorted=all(ents[i] <= ents[i + 1] for i in range(len(ents) - 1)), table_is_permutation=sorted(e[1] for e in ents) == list(range(d.m3)),
           table_sample_ok=all(ents[i][0] == o.pubkey(ents[i][1] + 1)[0].to_bytes(32, "big")[16:22] for i in rnd.sample(range(d.m3), 20)))
# planted key in the 38th window: 2^64 + 37*2^45 + offset
rnd = random.Random(4)
key = (1 << 64) + 37 * (1 << 45) + rnd.randrange(1 << 45)
pub = o.pubkey(key)
t0 = time.time(); got = kh.bsgs_search(pub, 1 << 64, 1 << 65); dt = time.time() - t0
st = kh.stats(reset=True)
res.update(planted_found=(got == key), time_to_find_s=dt, find_giant_steps=st["points"], find_walk_ms=st["walk_ms"], find_tier1_pos=st["tier1_positives"])
# key in the last window: the whole 2^64 range is swept (2^32 giant steps at k=512)
key2 = (1 << 65) - 12345
t0 = time.time(); got2 = kh.bsgs_search(o.pubkey(key2), 1 << 64, 1 << 65); dt2 = time.time() - t0
st = kh.stats(reset=True)
res.update(last_window_found=(got2 == key2), full_sweep_s=dt2, sweep_giant_steps=st["points"], sweep_walk_ms=st["walk_ms"], sweep_aux_ms=st["aux_ms"],
           sweep_tier1_pos=st["tier1_positives"], giant_steps_per_s=st["points"] / (st["walk_ms"] * 1e-3),
           keys_per_s=st["points"] / (st["walk_ms"] * 1e-3) * 2 * d.m)
print(json.dumps(res))
