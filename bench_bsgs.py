#!/usr/bin/env python3
"""bench_bsgs.py — C4 (BASELINE.json configs[3]): `-m bsgs -k 512`, n = 2^44, one public key in a 2^64-wide
range, bP table + 3-tier bloom resident in HBM.  Same JSON-line contract as bench.py (which stays on the
headline C2 workload); a step is one sweep of 2^15 windows (2^28 giant steps, 2^60 keys).

  python bench_bsgs.py [--k 512] [--steps 16] [--warmup 3] [--cpu-k 16] [--no-cpu-baseline]

Reports: table build time, giant steps/s and keys/s (= steps/s x 2m), time-to-find of the planted key
(2^64 + 37*2^45 + offset), roofline of kh_giant_kernel against the measured HBM bandwidth
(it is bound by random 32-byte sector reads of the 7.7 GB tier-1 bloom), and the unmodified reference
(`keyhunt -m bsgs -t <all cores>`) on the same host with a smaller -k (its k=512 table build alone takes
>= 14 min); giant steps/s is the k-independent figure to compare.
"""
import argparse
import json
import os
import random
import re
import shutil
import signal
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from bench import ClockSampler, cpu_model, log, ref_binary  # noqa: E402

GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
N44 = 1 << 44


def reference_bsgs_rate(k, seconds=25):
    """giant steps/s of the reference: build with -k, then read its own 'Total N keys in S seconds' lines"""
    exe = ref_binary()
    if exe is None:
        return None
    d = tempfile.mkdtemp(prefix="khref_bsgs_")
    try:
        # a valid public key that is NOT in the searched range (key 1 = G): the sweep never ends early
        open(os.path.join(d, "p.txt"), "w").write("0279be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798\n")
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        p = subprocess.Popen([exe, "-m", "bsgs", "-f", "p.txt", "-k", str(k), "-r", "10000000000000000:20000000000000000",
                              "-t", str(cores), "-q", "-s", "5", "-M"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                             text=True, start_new_session=True)
        build_s, last = None, None
        deadline = None
        for line in p.stdout:
            m = re.search(r"Total (\d+) keys in (\d+) seconds", line)
            if m:
                if build_s is None:
                    build_s = time.perf_counter() - t0 - int(m.group(2))
                    deadline = time.perf_counter() + seconds
                last = (int(m.group(1)), int(m.group(2)))
            if deadline and time.perf_counter() > deadline:
                break
        os.killpg(p.pid, signal.SIGKILL)
        p.wait()
        if not last:
            return None
        m_cpu = (1 << 22) * k
        keys_s = last[0] / last[1]
        return {"keys_per_s": keys_s, "giant_steps_per_s": keys_s / (2 * m_cpu), "k": k, "m": m_cpu, "cores": cores,
                "build_s": build_s, "sample_s": last[1]}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=512)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-k", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    import keyhunt_b200 as K
    kh = K.KeyHunt(0)
    info = kh.device_info()
    t0 = time.perf_counter()
    kh.bsgs_build(N44, args.k)
    build_wall = time.perf_counter() - t0
    st = kh.stats(reset=True)
    d = kh.bsgs_describe()
    build = {"wall_s": build_wall, "baby_walk_ms": st["walk_ms"], "sort_ms": st["aux_ms"], "baby_steps_per_s": d.m / (st["walk_ms"] * 1e-3),
             "tier1_GB": d.tier[0].bytes * 256 / 1e9, "m": d.m, "m2": d.m2, "m3": d.m3, "launches": st["walk_launches"] + st["other_launches"]}
    win = 2 * N44                      # keys per window
    W = 1 << 15                        # windows per step
    start = 1 << 64
    # planted key (found in window 37) — time to find from the range start
    rnd = random.Random(4)
    key = start + 37 * (1 << 45) + rnd.randrange(1 << 45)
    pub = kh.derive([key])[0]
    t0 = time.perf_counter()
    got = kh.bsgs_search((pub.pub_x, pub.pub_y), start, 1 << 65)
    t_find = time.perf_counter() - t0
    kh.stats(reset=True)
    # timed sweeps: a key outside the range, W windows per step (every step a new sub-range)
    for s in range(args.warmup):
        kh.bsgs_search((GX, GY), start + s * W * win, start + (s + 1) * W * win)
    kh.stats(reset=True)
    clocks = ClockSampler(0)
    clocks.start()
    t0 = time.perf_counter()
    for s in range(args.steps):
        a = start + (args.warmup + s) * W * win
        kh.bsgs_search((GX, GY), a, a + W * win)
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    st = kh.stats(reset=True)
    dev_ms = st["walk_ms"] + st["setup_ms"] + st["aux_ms"]
    steps_total = st["points"]
    gs = steps_total / (st["walk_ms"] * 1e-3)
    mp = {}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = mp.get("hbm_gbs", 6650.0)
    cpu = None
    if not args.no_cpu_baseline:
        try:
            r = reference_bsgs_rate(args.cpu_k)
            if r:
                cpu = {"value": r["keys_per_s"] / 1e15, "unit": "Pkeys/s", "cores": r["cores"], "kind": "reference",
                       "sample": "keyhunt -m bsgs -k %d (m=2^%d) -t %d, %d s of its own stats line after a %.0f s table build, %s"
                                 % (r["k"], r["m"].bit_length() - 1, r["cores"], r["sample_s"], r["build_s"] or -1, cpu_model()),
                       "giant_steps_per_s": r["giant_steps_per_s"], "note": "keys/s scales with the table size m; giant steps/s is the k-independent figure"}
        except Exception as e:
            cpu = {"value": None, "unit": "Pkeys/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}
    keys_s = steps_total * 2 * d.m / (dev_ms * 1e-3)
    line = {
        "metric": "Pkeys/s (c4 bsgs -k %d)" % args.k, "value": keys_s / 1e15, "unit": "Pkeys/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": "C4 bsgs -k %d, n=2^44 (m=2^%d), 1 public key, range [2^64, 2^65), step = 2^15 windows = 2^28 giant steps"
                               % (args.k, d.m.bit_length() - 1), "gpu": info["name"],
                   "l2_note": "tier-1 bloom (%.1f GB) and its prefix bitmap (64 GB) are far larger than L2; every probe is a random HBM sector" % build["tier1_GB"]},
        "clocks": clk, "build": build,
        "giant_steps_per_s": gs, "wall_ms_per_step": wall * 1e3 / args.steps,
        "planted": {"found": got == key, "time_to_find_s": t_find},
        "e2e": {"value": steps_total * 2 * d.m / wall / 1e15, "unit": "Pkeys/s", "h2d_bytes_per_step": 128, "d2h_bytes_per_step": 48, "steps": args.steps},
        "gpu_launches": st["walk_launches"] + st["other_launches"], "tier1_positives": st["tier1_positives"],
        "roofline": {"bound": "hbm", "achieved": gs * 64 / 1e9, "peak": hbm, "unit": "GB/s", "frac": gs * 64 / 1e9 / hbm,
                     "traffic": 161.0 * (st["points"] / max(1, st["walk_launches"])),
                     "kernel": "kh_giant_kernel", "bytes_per_giant_step": 64,
                     "note": "algorithmic 64 B/step (SURVEY §8d: 2 random 32-B sectors; here 16 B + 16 B of prefix-product scratch and one "
                             "32-B sector of the baby-point prefix bitmap that answers for the tier-1 bloom); ncu measures 145 B read + "
                             "16 B written per step: the single random probe into the 64 GB bitmap costs ~4 sectors (a 64-B DRAM atom plus "
                             "page-table reads), DRAM 38.7 % of peak on a purely random pattern, FMA-heavy pipe 69 % busy "
                             "(profiles/r01_giant_prefilter_ncu_metrics.csv; without the bitmap: 243 B per step, r01_giant_ncu_metrics.csv)",
                     "int_ops_per_step": 920, "int_tiops": gs * 920 / 1e12},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    kh.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
