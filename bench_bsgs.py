#!/usr/bin/env python3
"""bench_bsgs.py — C4 (BASELINE.json configs[3]) alone: `-m bsgs -k 512`, n = 2^44, one public key in [2^64, 2^65), bP table +
3-tier bloom resident in HBM.  The same measurement is part of the one line `python bench.py` prints (`workloads.c4`); this
script prints it on its own, with the driver's JSON-line contract.

  python bench_bsgs.py [--k 512] [--steps 16] [--warmup 3] [--cpu-k 16] [--no-cpu-baseline]

A step is one sweep of 2^15 windows (2^28 giant steps, 2^60 keys).  Reports table build time, giant steps/s through
kh_bsgs_search and kernel-only, time-to-find of the planted key, the roofline of kh_giant_kernel, and the unmodified
reference (`keyhunt -m bsgs -t <all cores>`) on the same host at a smaller -k in the same unit (giant steps/s do not depend on
k; its k = 512 table build alone takes >= 14 min).
"""
import argparse
import json
import sys

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=512)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-k", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    args = argparse.Namespace(steps_per_launch=0)
    env = bench.Env(args)
    env.peaks = env.kh.int_peak()
    r = bench.run_c4(env, a.k, a.steps, max(3, a.warmup), a.cpu_k, 0 if a.no_cpu_baseline else 14)
    line = {"metric": r["metric"], "value": r["value"], "unit": r["unit"], "n_gpus": 1, "steps": r["steps"], "warmup": r["warmup"],
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic"}
    line.update({k: v for k, v in r.items() if k not in line})
    print(json.dumps(line))
    env.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
