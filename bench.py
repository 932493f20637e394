#!/usr/bin/env python3
"""bench.py — measures BASELINE.json's metric (Mkeys/s of the bounded key-range search) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c5btc|c5eth] [--impl reference]

A *step* is one pass of the hot path over one batch of synthetic input: a contiguous sub-range of
2^32 keys (per GPU) scanned against the workload's target set.  Default workload = C2 (BASELINE.json
configs[1]): rmd160 mode, compress+uncompress, 1,024 synthetic hash160 targets (24 planted), and the
default K=16 steps cover exactly the 2^36-key range of that config.  Keys/s = points/s for `-l both`
(the reference counts 1024 per batch, keyhunt.cpp:2876-2891); for compress-only workloads the line also
carries the reference's "displayed" figure (x2).

value  : whole-job throughput with the target set already resident in HBM, timed on the device
         (CUDA events on the library's stream around every kernel of the step), max over ranks.
e2e    : the same metric through the reference-facing C ABI with HOST buffers: every step uploads the
         target records from pinned host memory (kh_set_targets: H2D + on-device bloom build), scans
         (kh_scan) and reads the hits back (kh_poll_hits: D2H); wall clock around the loop.
--impl reference : the unmodified reference CPU tool (oracle/_ref/keyhunt*, built by oracle/Makefile from
         /root/reference) with all host threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import random
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ORDER = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
STEP_POINTS = 1 << 32

# algorithmic 32-bit integer ops per point (the constants the roofline uses)
WORKLOADS = {
    # name: (mode, crypto, search, range start, n targets, planted, ops/point, reference flags, displayed multiplier)
    # ops/point = the SURVEY §8(d) constant minus 162 per hashed record: the exact prefix bitmap (emit.cuh prefilter_pass, ~8 ops)
    # answers for the 170-op bloom_check of a non-member, so that work is no longer done and must not be counted as achieved
    "c1": dict(desc="C1 address compress, tests/1to32 puzzle targets", mode="address", crypto="btc", search="compress",
               start=0x1, n_targets=32, ops=5800 - 2 * 162, disp=2),
    "c2": dict(desc="C2 rmd160 -l both, 1024 hash160 targets (24 planted), 2^36 keys from 0x2000000000000000",
               mode="rmd160", crypto="btc", search="both", start=0x2000000000000000, n_targets=1024, planted=24, ops=9950 - 3 * 162, disp=1,
               alu_ops=6830),   # ncu (profiles/r01_prefilter_both_ncu_sections.txt): 42.46 G warp instructions per 2^27 points, ALU share = (86.5 % x 0.5/clk) / 64.1 % issue = 67.5 % -> 6,830 ALU thread-ops per point
    "c3": dict(desc="C3 xpoint, 10^6 x-coordinates (32 planted), 2^36 keys from 0x4000000000000000",
               mode="xpoint", crypto="btc", search="compress", start=0x4000000000000000, n_targets=1000000, planted=32, ops=900 - 162, disp=1),
    "c5btc": dict(desc="C5 address BTC compress, 1024 targets (16 planted), from 0x10000000000",
                  mode="address", crypto="btc", search="compress", start=0x10000000000, n_targets=1024, planted=16, ops=5800 - 2 * 162, disp=2),
    "c5eth": dict(desc="C5 address ETH, 1024 targets (16 planted), from 0x10000000000",
                  mode="address", crypto="eth", search="compress", start=0x10000000000, n_targets=1024, planted=16, ops=5930 - 162, disp=1),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------------
def planted_indices(seed, n_points_total, count):
    """first key, last key, and uniformly drawn keys inside the scanned range"""
    rnd = random.Random(seed)
    idx = {0, n_points_total - 1}
    while len(idx) < count:
        idx.add(rnd.randrange(n_points_total))
    return sorted(idx)


def make_targets(kh, K, wl, seed, n_points_total):
    """(records20 bytes, {record: (index, key)}) — planted records are derived on the device from their
    private keys (kh_derive), decoys are random bytes."""
    w = WORKLOADS[wl]
    rnd = random.Random(seed * 7919 + 1)
    planted = {}
    if wl == "c1":
        recs = K.parse_targets(open(os.path.join(ROOT, "tests", "golden", "1to32.txt")), K.MODE_ADDRESS)
        return recs, planted
    idxs = planted_indices(seed, n_points_total, w["planted"])
    infos = kh.derive([w["start"] + i for i in idxs])
    recs = []
    for j, (i, info) in enumerate(zip(idxs, infos)):
        if w["mode"] == "xpoint":
            r = info.pub_x.to_bytes(32, "big")[:20]
        elif w["crypto"] == "eth":
            r = info.eth
        elif w["search"] == "both":
            r = info.h160_uncomp if j % 2 else info.h160_comp
        elif w["search"] == "uncompress":
            r = info.h160_uncomp
        else:
            r = info.h160_comp
        planted[r] = (i, w["start"] + i)
        recs.append(r)
    while len(recs) < w["n_targets"]:
        recs.append(rnd.randbytes(20))
    recs.sort()      # the reference's boundary hands over the sorted addressTable (_sort, keyhunt.cpp:1361)
    return b"".join(recs), planted


def kh_modes(K, wl):
    w = WORKLOADS[wl]
    mode = {"address": K.MODE_ADDRESS, "rmd160": K.MODE_RMD160, "xpoint": K.MODE_XPOINT}[w["mode"]]
    crypto = K.CRYPTO_ETH if w["crypto"] == "eth" else K.CRYPTO_BTC
    search = {"compress": K.SEARCH_COMPRESS, "uncompress": K.SEARCH_UNCOMPRESS, "both": K.SEARCH_BOTH}[w["search"]]
    return mode, crypto, search


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# reference CPU tool
# ------------------------------------------------------------------------------------------------------
def ref_binary():
    flags = open("/proc/cpuinfo").read()
    v3 = os.path.join(ROOT, "oracle", "_ref", "keyhunt_v3")
    v2 = os.path.join(ROOT, "oracle", "_ref", "keyhunt")
    if os.path.exists(v3) and all(f in flags for f in (" avx2", " bmi2", " adx")):
        return v3
    return v2 if os.path.exists(v2) else None


def cpu_model():
    for ln in open("/proc/cpuinfo"):
        if ln.startswith("model name"):
            return ln.split(":", 1)[1].strip()
    return "unknown"


def run_reference(wl, records20, start, n_points, threads, chunk=1 << 20):
    """one run of the unmodified reference on [start, start+n_points); returns (seconds, keys found)"""
    w = WORKLOADS[wl]
    exe = ref_binary()
    if exe is None:
        raise RuntimeError("oracle/_ref/keyhunt is missing (build it with `make -C oracle ref` where /root/reference exists)")
    d = tempfile.mkdtemp(prefix="khref_")
    try:
        fn = os.path.join(d, "targets.txt")
        with open(fn, "w") as f:
            for i in range(0, len(records20), 20):
                r = records20[i:i + 20]
                if w["mode"] == "xpoint":
                    f.write(r.hex() + "00" * 12 + "\n")      # 64-hex X value; only the first 20 bytes are compared
                elif w["crypto"] == "eth":
                    f.write("0x" + r.hex() + "\n")
                else:
                    f.write(r.hex() + "\n")
        mode = "rmd160" if (w["mode"] in ("rmd160", "address") and w["crypto"] == "btc") else w["mode"]
        cmd = [exe, "-m", mode, "-f", fn, "-r", "%x:%x" % (start, start + n_points), "-n", hex(chunk), "-t", str(threads),
               "-q", "-s", "0"]
        if w["mode"] != "xpoint":
            cmd += ["-l", w["search"]]
        if w["crypto"] == "eth":
            cmd += ["-c", "eth"]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        dt = time.perf_counter() - t0
        if r.returncode != 0 or "End" not in r.stdout:
            raise RuntimeError("reference run failed: %s\n%s" % (" ".join(cmd), r.stdout[-2000:]))
        keys = []
        kf = os.path.join(d, "KEYFOUNDKEYFOUND.txt")
        if os.path.exists(kf):
            for ln in open(kf):
                if ln.startswith("Private Key:"):
                    keys.append(int(ln.split(":")[1].strip(), 16))
        return dt, sorted(keys)
    finally:
        shutil.rmtree(d, ignore_errors=True)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload
    w = WORKLOADS[wl]
    if ref_binary() is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/keyhunt not built (needs /root/reference at build time)"}))
        return 0
    cores = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    per_step_s = min(20.0, 150.0 / max(1, K + W))
    chunks = max(1, int(per_step_s * 0.6e6 / (1 << 20)))           # ~0.6 Mpoints/s per thread on -l both
    n_points = cores * chunks * (1 << 20)
    rnd = random.Random(2)
    if wl == "c1":
        import keyhunt_b200 as KH
        recs = KH.parse_targets(open(os.path.join(ROOT, "tests", "golden", "1to32.txt")), KH.MODE_ADDRESS)
    else:
        recs = b"".join(rnd.randbytes(20) for _ in range(w["n_targets"]))
    times = []
    for s in range(W + K):
        dt, _ = run_reference(wl, recs, w["start"] + s * n_points, n_points, cores)
        if s >= W:
            times.append(dt)
        log("[reference] step %d: %.2f s for %d keys" % (s, dt, n_points))
    tot = sum(times)
    val = K * n_points / tot / 1e6
    line = {"impl": "reference", "metric": "Mkeys/s (%s, points/s)" % wl, "value": val, "unit": "Mkeys/s", "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * tot / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": w["desc"], "sample_keys_per_step": n_points, "binary": os.path.basename(ref_binary()),
                       "flags": "-t %d -n 0x100000 -q -s 0" % cores},
            "cpu_baseline": {"value": val, "unit": "Mkeys/s", "cores": cores, "kind": "reference",
                             "sample": "%d steps x %d keys of the %s range, %s" % (K, n_points, wl, cpu_model())},
            "e2e": {"value": val, "unit": "Mkeys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--steps-per-launch", type=int, default=0, help="library option (profiling runs use 1 for short kernels)")
    ap.add_argument("--step-points-log2", type=int, default=32, help="keys per step = 2^this (profiling runs use less)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # not under torchrun: relaunch one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.warmup < 3:
        log("[bench] warm-up raised to 3 (timing rules)")
        args.warmup = 3

    # stdout must carry exactly one JSON line: NCCL's version/debug banner goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import keyhunt_b200 as K
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = args.workload
    w = WORKLOADS[wl]
    Ksteps, W = args.steps, args.warmup
    kh = K.KeyHunt(local)
    if args.steps_per_launch:
        kh.set_option("steps_per_launch", args.steps_per_launch)
    global STEP_POINTS
    STEP_POINTS = 1 << args.step_points_log2
    info = kh.device_info()
    mode, crypto, search = kh_modes(K, wl)

    # global range: N ranks x K steps x 2^32 keys; rank r owns steps [r*K, (r+1)*K) (contiguous shard, no collective)
    total_points = world * Ksteps * STEP_POINTS
    records, planted = make_targets(kh, K, wl, seed=2, n_points_total=total_points)
    rec_host = torch.frombuffer(bytearray(records), dtype=torch.uint8).pin_memory()     # pinned host copy of the targets
    rec_ptr = rec_host.data_ptr()
    import ctypes as C
    rec_c = (C.c_char * len(records)).from_address(rec_ptr)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    shard_start = w["start"] + rank * Ksteps * STEP_POINTS
    kh.set_targets(mode, records, crypto=crypto, search=search)
    peaks = kh.int_peak() if rank == 0 else None

    # ---- warm-up (sub-ranges below the range start; their hits are discarded) ------------------------
    for s in range(W):
        kh.scan(w["start"] - (s + 1) * STEP_POINTS if w["start"] > (W + 1) * STEP_POINTS else w["start"] + (total_points + s * STEP_POINTS), STEP_POINTS)
    kh.poll_hits()
    kh.stats(reset=True)

    # ---- timed: device-resident ----------------------------------------------------------------------
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    t0 = time.perf_counter()
    for s in range(Ksteps):
        kh.scan(shard_start + s * STEP_POINTS, STEP_POINTS)
    hits = kh.poll_hits()
    barrier()
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    st = kh.stats(reset=True)
    dev_ms = st["walk_ms"] + st["setup_ms"] + st["aux_ms"]

    # ---- timed: end to end through the C ABI with host buffers ---------------------------------------
    e2e_steps = max(1, min(args.e2e_steps, Ksteps))
    barrier()
    t0 = time.perf_counter()
    e2e_hits = []
    for s in range(e2e_steps):
        kh._ck(kh._lib.kh_set_targets(kh._h, mode, crypto, search, rec_c, len(records) // 20, None, None))   # H2D from pinned memory
        kh.scan(shard_start + s * STEP_POINTS, STEP_POINTS)
        e2e_hits += kh.poll_hits()                                                                           # D2H
    barrier()
    e2e_wall = time.perf_counter() - t0
    st_e2e = kh.stats(reset=True)

    # ---- reduce over ranks (max time; hits gathered to rank 0) ---------------------------------------
    found = sorted((h.index, h.key, h.matched.hex(), h.kind) for h in hits)   # index is relative to its step's start
    times = torch.tensor([dev_ms, wall * 1e3, e2e_wall * 1e3], dtype=torch.float64, device="cuda")
    launches = torch.tensor([st["walk_launches"] + st["other_launches"]], dtype=torch.int64, device="cuda")
    all_found = [found]
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
        gathered = [None] * world
        dist.all_gather_object(gathered, found)
        all_found = gathered
    dev_ms_max, wall_ms_max, e2e_ms_max = [float(x) for x in times.tolist()]
    if rank != 0:
        kh.close()
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- checks: every planted key found, nothing else -----------------------------------------------
    flat = [f for part in all_found for f in part]
    got_keys = sorted(f[1] for f in flat)
    if wl == "c1":
        ok = None
    else:
        want_keys = sorted(k for (_, k) in planted.values())
        ok = (got_keys == want_keys)
        if not ok:
            log("[bench] HIT MISMATCH: got %d want %d" % (len(got_keys), len(want_keys)))

    # ---- CPU baseline: the unmodified reference on a bounded sample of the same workload --------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:   # rank 0 at N=1 only
        try:
            cores = os.cpu_count() or 1
            chunks = 16 if w["ops"] > 3000 else 48
            n_cpu = cores * chunks * (1 << 20)
            # sample = the head of the range (contains planted key index 0) -> also a hit-parity check
            dt, ref_keys = run_reference(wl, records, w["start"], n_cpu, cores)
            mine = sorted(f[1] for f in flat if w["start"] <= f[1] < w["start"] + n_cpu) if wl != "c1" else None
            # the reference tests its range cursor outside the mutex (keyhunt.cpp:3314), so racing threads may scan a few
            # chunks past the end: compare inside the sample only
            ref_keys = [k for k in ref_keys if w["start"] <= k < w["start"] + n_cpu]
            cpu = {"value": n_cpu / dt / 1e6, "unit": "Mkeys/s", "cores": cores, "kind": "reference",
                   "sample": "first %d keys of the %s range, %s -t %d -n 0x100000, %s, %.1f s wall" %
                             (n_cpu, wl, os.path.basename(ref_binary()), cores, cpu_model(), dt),
                   "hits_equal_gpu": (mine == ref_keys) if mine is not None else None, "hits": len(ref_keys)}
        except Exception as e:  # the bench line must still come out
            cpu = {"value": None, "unit": "Mkeys/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}

    points = world * Ksteps * STEP_POINTS
    value = points / (dev_ms_max * 1e-3) / 1e6
    # roofline of the dominant kernel (kh_scan_kernel): algorithmic int ops per launch / mean launch duration
    pts_per_launch = Ksteps * STEP_POINTS / max(1, st["walk_launches"])
    launch_ms = st["walk_ms"] / max(1, st["walk_launches"])
    achieved = pts_per_launch * w["ops"] / (launch_ms * 1e-3) / 1e12
    peak = peaks["lop3_imad_mix"] / 1e12
    mp = {}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    nominal = info["sm_count"] * 64 * (clk["sm_mhz"] or 1965.0) * 1e6 / 1e12
    line = {
        "metric": "Mkeys/s (%s, points/s)" % wl, "value": value, "unit": "Mkeys/s", "n_gpus": world, "steps": Ksteps, "warmup": W,
        "ms_per_step": dev_ms_max / Ksteps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": w["desc"], "keys_per_step_per_gpu": STEP_POINTS, "keys_total": points, "targets": len(records) // 20,
                   "l2_note": "inputs larger than L2: every step walks 2^32 new keys; per-thread scratch (1.2 GB) streams through HBM",
                   "displayed_keys_multiplier": w["disp"], "walker_threads": st["walker_threads"], "gpu": info["name"]},
        "clocks": clk,
        "e2e": {"value": world * e2e_steps * STEP_POINTS / (e2e_ms_max * 1e-3) / 1e6, "unit": "Mkeys/s",
                "h2d_bytes_per_step": len(records) + 64, "d2h_bytes_per_step": 8 + 160 * max(1, len(e2e_hits)) // max(1, e2e_steps),
                "steps": e2e_steps, "launches": st_e2e["walk_launches"] + st_e2e["other_launches"]},
        "gpu_launches": int(launches.item()),
        "displayed_keys_value": value * w["disp"],   # the reference multiplies by 2 for -l compress (keyhunt.cpp:2889-2891)
        "wall_ms_per_step": wall_ms_max / Ksteps,
        "hits": {"found": len(got_keys), "all_planted_found_and_nothing_else": ok},
        "roofline": {"bound": "int", "achieved": achieved, "peak": peak, "unit": "Tiop/s", "frac": achieved / peak,
                     # DRAM bytes per launch: ncu --set full measured 39.23 GB for a 1.2416 G-point launch of this kernel
                     # (profiles/r01_v1_scan_both_ncu_full_summary.txt) = 31.6 B/point = the algorithmic scratch write+read
                     "traffic": 31.6 * pts_per_launch, "traffic_unit": "bytes of DRAM read+write per launch (ncu-measured 31.6 B/point x points per launch)",
                     "kernel": "kh_scan_kernel", "ops_per_point": w["ops"], "launch_ms": launch_ms,
                     "peak_source": "measured live: kh_int_peak LOP3+IMAD dual-pipe rate; ALU pipe alone %.2f, IMAD %.2f, IMAD.WIDE %.2f Tiop/s"
                                    % (peaks["lop3"] / 1e12, peaks["imad"] / 1e12, peaks["imad_wide"] / 1e12),
                     "frac_of_nominal_64_lanes": achieved / nominal, "nominal_peak": nominal,
                     "hbm_gbs_scratch": 32.0 * value * 1e6 / 1e9, "hbm_peak_gbs": mp.get("hbm_gbs"),
                     # the pipe that actually binds (ncu: ALU 84 % busy, top stall math_pipe_throttle): ALU ops per point x points/s
                     # against the live-measured ALU-only rate
                     "binding_pipe": ({"pipe": "alu", "ops_per_point": w["alu_ops"],
                                       "achieved": pts_per_launch * w["alu_ops"] / (launch_ms * 1e-3) / 1e12, "peak": peaks["lop3"] / 1e12,
                                       "frac": pts_per_launch * w["alu_ops"] / (launch_ms * 1e-3) / peaks["lop3"], "unit": "Tiop/s"}
                                      if "alu_ops" in w else None)},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    kh.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
