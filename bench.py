#!/usr/bin/env python3
"""bench.py — measures BASELINE.json's metric (Mkeys/s of the bounded key-range search) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c5btc|c5eth] [--impl reference]
                  [--no-side-workloads] [--no-strong] [--no-cpu-baseline]

A *step* is one pass of the hot path over one batch of synthetic input: a contiguous sub-range of 2^32 keys (per GPU)
scanned against the workload's target set.  Default workload = C2 (BASELINE.json configs[1]): rmd160 mode,
compress+uncompress, 1,024 synthetic hash160 targets (24 planted); K=16 steps cover exactly its 2^36-key range.
Keys/s = points/s for `-l both` (the reference counts 1024 per batch, keyhunt.cpp:2876-2891); for compress-only workloads
the line also carries the reference's "displayed" figure (x2).

The ONE JSON line (rank 0, stdout):
  value      whole-job throughput of the main workload with the target set resident in HBM, device-timed (CUDA events on
             the library's stream around every kernel of the step), max over ranks.
  e2e        the same metric through the reference-facing C ABI with HOST buffers: every step uploads the target records
             from pinned host memory (kh_set_targets: H2D + on-device bloom build), scans (kh_scan) and reads the hits back
             (kh_poll_hits: D2H); wall clock around the loop.
  roofline   dominant kernel against the live-measured integer peaks (kh_int_peak) — see DESIGN.md §4.
  cpu_baseline  the unmodified reference CPU tool on all host cores on the head of the same range (N=1 only).
  workloads  (N=1) the OTHER BASELINE configs, each a few timed steps with value / e2e / roofline / planted-hit check /
             reference CPU figure: c1, c3, c5btc, c5eth (scan modes) and c4 (bsgs -k 512, giant steps/s next to the
             reference's giant steps/s, which do not depend on k).
  strong     strong scaling: a FIXED range cut over the N ranks (keyhunt_b200.sharding), clock around target upload +
             set-up + scan + NCCL hit gather: C5 (2^37 keys, ETH and BTC-compress hit sets separately) and C4 (windows of a
             2^66-wide range dealt over the ranks, tables built on every GPU, build time reported).
  peaks      the measured pipe rates the roofline uses (also for the driver to copy next to MEASURED_PEAKS.json).

--impl reference : the unmodified reference CPU tool (oracle/_ref/keyhunt*, built by oracle/Makefile from /root/reference)
  with all host threads on a bounded sample of the same workload; its rate is read from the reference's OWN statistics
  lines ("Total N keys in S seconds", keyhunt.cpp:2907) in steady state — between the thread ramp-up of the first seconds
  and the drain of the last chunks — so that neither process start-up nor the 1-second polling of its main loop
  (keyhunt.cpp:2840) is charged to it; the plain wall-clock figure is reported beside it.
"""
import argparse
import ctypes as C
import json
import os
import random
import re
import shutil
import signal
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ORDER = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
STEP_POINTS = 1 << 32

# algorithmic 32-bit integer ops per point (the constants the roofline uses).
# ops = the SURVEY constant minus 162 per hashed record: the exact prefix bitmap (emit.cuh prefilter_pass, ~8 ops) answers
# for the 170-op bloom_check of a non-member, so that work is no longer done and must not be counted as achieved.  Likewise the
# message schedule of the second SHA-256 block of the uncompressed key (48 words x 10 ops in SURVEY's count) is looked up in a
# 256-row table since round 2 (hash.cuh KH_SHA_UNC2_TAB) and is not counted: C2 ops = 9950 - 3*162 - 480; and the ~270 lane operations that
# Keccak's peeled first / last round no longer execute (zero input lanes, unused output lanes: hash.cuh KH_KECCAK_PEEL) are not counted for ETH.
# cpu_rate = expected Mkeys/s per host thread of the reference (only used to size its bounded sample).
WORKLOADS = {
    "c1": dict(desc="C1 address compress, tests/1to32 puzzle targets", mode="address", crypto="btc", search="compress",
               start=0x1, n_targets=32, ops=5800 - 2 * 162, survey_ops=5800, disp=2, cpu_rate=2.4, binding="alu", alu_ops=3888),
    "c2": dict(desc="C2 rmd160 -l both, 1024 hash160 targets (24 planted), 2^36 keys from 0x2000000000000000",
               mode="rmd160", crypto="btc", search="both", start=0x2000000000000000, n_targets=1024, planted=24, ops=9950 - 3 * 162 - 480, survey_ops=9950, disp=1,
               cpu_rate=1.4, binding="alu",
               alu_ops=6625),   # executed ALU-pipe thread instructions per point, from the ncu source page (profiles/r02_both_opmix.txt: 40.21 G warp instructions per 2^27 points, 69.1 % of them SHF/LOP3/IADD3/LEA/...); likewise for the other kinds
    "c3": dict(desc="C3 xpoint, 10^6 x-coordinates (32 planted), 2^36 keys from 0x4000000000000000",
               mode="xpoint", crypto="btc", search="compress", start=0x4000000000000000, n_targets=1000000, planted=32, ops=900 - 162, survey_ops=900, disp=1,
               cpu_rate=4.8, binding="fma_heavy", wide_mults=244, dram_b_per_point=37.3),   # executed IMAD.WIDE per point (profiles/r02_xpoint_opmix.txt; 2.5 M + 1 S = 224 of them, the rest is bloom/bitmap index arithmetic)
    "c5btc": dict(desc="C5 address BTC compress, 1024 targets (16 planted), from 0x10000000000",
                  mode="address", crypto="btc", search="compress", start=0x10000000000, n_targets=1024, planted=16, ops=5800 - 2 * 162, survey_ops=5800, disp=2,
                  cpu_rate=2.4, binding="alu", alu_ops=3888),
    "c5eth": dict(desc="C5 address ETH, 1024 targets (16 planted), from 0x10000000000",
                  mode="address", crypto="eth", search="compress", start=0x10000000000, n_targets=1024, planted=16, ops=5930 - 162 - 270, survey_ops=5930, disp=1,
                  cpu_rate=2.1, binding="alu", alu_ops=5268),
}
N44 = 1 << 44          # C4: -n 2^44 -k 512 -> m = 2^31 baby steps


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


def steady_rate(stats, mult=1):
    """keys/s in steady state from the reference's own "Total N keys in S seconds" lines [(N, S), ...]: the increments of the
    first two seconds (thread ramp-up) and of the last two (threads running dry) are left out.  None if the run was too short."""
    pts = sorted(set((s, n) for n, s in stats))
    if len(pts) < 6:
        return None
    a, b = pts[2], pts[-3]
    if b[0] <= a[0]:
        return None
    return (b[1] - a[1]) / float(b[0] - a[0]) / mult


def ref_sample_points(wl, cores, seconds):
    """sample size for about `seconds` of reference work: whole 2^20-key chunks, the same number per thread"""
    chunks = max(2, int(seconds * WORKLOADS[wl]["cpu_rate"] * 1e6 / (1 << 20)))
    return cores * chunks * (1 << 20)


def run_reference(wl, records20, start, n_points, threads, chunk=1 << 20):
    """one run of the unmodified reference on [start, start+n_points) -> dict(wall_s, keys found, steady keys/s or None)"""
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
               "-q", "-s", "1"]
        mult = 1
        if w["mode"] != "xpoint":
            cmd += ["-l", w["search"]]
            mult = 2 if w["search"] == "compress" else 1     # the statistics line doubles the count for -l compress (keyhunt.cpp:2889)
        if w["crypto"] == "eth":
            cmd += ["-c", "eth"]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, errors="replace")
        dt = time.perf_counter() - t0
        if r.returncode != 0 or "End" not in r.stdout:
            raise RuntimeError("reference run failed: %s\n%s" % (" ".join(cmd), r.stdout[-2000:]))
        stats = [(int(m.group(1)), int(m.group(2))) for m in re.finditer(r"Total (\d+) keys in (\d+) seconds", r.stdout)]
        keys = []
        kf = os.path.join(d, "KEYFOUNDKEYFOUND.txt")
        if os.path.exists(kf):
            for ln in open(kf):
                if ln.startswith("Private Key:"):
                    keys.append(int(ln.split(":")[1].strip(), 16))
        return {"wall_s": dt, "keys": sorted(keys), "steady": steady_rate(stats, mult), "flags": " ".join(cmd[cmd.index("-n"):])}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def reference_bsgs_rate(k, seconds=14):
    """giant steps/s of the reference in steady state: build with -k, then its own 'Total N keys in S seconds' lines"""
    exe = ref_binary()
    if exe is None:
        return None
    d = tempfile.mkdtemp(prefix="khref_bsgs_")
    try:
        # a valid public key that is NOT in the searched range (key 1 = G): the sweep never ends early
        open(os.path.join(d, "p.txt"), "w").write("0279be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798\n")
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        p = subprocess.Popen(["stdbuf", "-oL", exe, "-m", "bsgs", "-f", "p.txt", "-k", str(k), "-r", "10000000000000000:20000000000000000",
                              "-t", str(cores), "-q", "-s", "1", "-M"] if shutil.which("stdbuf") else
                             [exe, "-m", "bsgs", "-f", "p.txt", "-k", str(k), "-r", "10000000000000000:20000000000000000",
                              "-t", str(cores), "-q", "-s", "1", "-M"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                             text=True, errors="replace", start_new_session=True)
        build_s, stats, deadline = None, [], None
        for line in p.stdout:
            for m in re.finditer(r"Total (\d+) keys in (\d+) seconds", line):
                if build_s is None:
                    build_s = time.perf_counter() - t0 - int(m.group(2))
                    deadline = time.perf_counter() + seconds
                stats.append((int(m.group(1)), int(m.group(2))))
            if deadline and time.perf_counter() > deadline:
                break
        os.killpg(p.pid, signal.SIGKILL)
        p.wait()
        if len(stats) < 2:
            return None
        m_cpu = (1 << 22) * k
        pts = sorted(set((s, n) for n, s in stats))
        a, b = pts[min(2, len(pts) - 2)], pts[-1]          # after the ramp-up; the sweep never drains (it is killed)
        keys_s = (b[1] - a[1]) / float(max(1, b[0] - a[0]))
        return {"keys_per_s": keys_s, "giant_steps_per_s": keys_s / (2 * m_cpu), "k": k, "m": m_cpu, "cores": cores,
                "build_s": build_s, "sample_s": b[0] - a[0]}
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
    per_step_s = max(8.0, min(16.0, 200.0 / max(1, K + W)))      # >= 8 s of work per process: enough statistics lines for the steady state
    n_points = ref_sample_points(wl, cores, per_step_s)
    rnd = random.Random(2)
    if wl == "c1":
        import keyhunt_b200 as KH
        recs = KH.parse_targets(open(os.path.join(ROOT, "tests", "golden", "1to32.txt")), KH.MODE_ADDRESS)
    else:
        recs = b"".join(rnd.randbytes(20) for _ in range(w["n_targets"]))
    secs, walls, flags = [], [], ""
    for s in range(W + K):
        r = run_reference(wl, recs, w["start"] + s * n_points, n_points, cores)
        flags = r["flags"]
        eff = n_points / r["steady"] if r["steady"] else r["wall_s"]
        if s >= W:
            secs.append(eff)
            walls.append(r["wall_s"])
        log("[reference] step %d: %d keys, steady state %.2f Mkeys/s (%.2f s of scanning), process wall %.2f s"
            % (s, n_points, n_points / eff / 1e6, eff, r["wall_s"]))
    tot = sum(secs)
    val = K * n_points / tot / 1e6
    line = {"impl": "reference", "metric": "Mkeys/s (%s, points/s)" % wl, "value": val, "unit": "Mkeys/s", "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * tot / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": w["desc"], "sample_keys_per_step": n_points, "binary": os.path.basename(ref_binary()),
                       "flags": flags,
                       "timing": "the reference's own statistics lines (Total N keys in S seconds), steady state: first 2 s (thread ramp-up) and last "
                                 "2 s (drain) of every process left out; each step is one process on its own sub-range"},
            "wall_value": K * n_points / sum(walls) / 1e6, "wall_ms_per_step": 1e3 * sum(walls) / K,
            "cpu_baseline": {"value": val, "unit": "Mkeys/s", "cores": cores, "kind": "reference",
                             "sample": "%d steps x %d keys of the %s range, %s" % (K, n_points, wl, cpu_model())},
            "e2e": {"value": val, "unit": "Mkeys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
class Env:
    """what every measurement needs: library, context, ranks, barrier"""

    def __init__(self, args):
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout must carry exactly one JSON line
        import torch
        import keyhunt_b200 as K
        self.torch, self.K, self.dist = torch, K, None
        if self.world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.kh = K.KeyHunt(self.local)
        if args.steps_per_launch:
            self.kh.set_option("steps_per_launch", args.steps_per_launch)
        self.info = self.kh.device_info()
        self.peaks = None
        self.pipe = None
        self.c4_build_s = None
        self.mp = {}
        try:
            self.mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def reduce(self, values, op="max"):
        """list of floats -> reduced over ranks on the device (NCCL)"""
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def gather(self, obj):
        if self.dist is None:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        self.kh.close()
        if self.dist is not None:
            self.dist.destroy_process_group()


def roofline(env, wl, value_pts_s, pts_per_launch, launch_ms, clk):
    """dominant kernel (kh_scan_kernel): algorithmic int ops per launch / mean launch duration against the live-measured peak"""
    w, peaks = WORKLOADS[wl], env.peaks
    achieved = pts_per_launch * w["ops"] / (launch_ms * 1e-3) / 1e12
    peak = peaks["lop3_imad_mix"] / 1e12
    nominal = env.info["sm_count"] * 64 * ((clk or {}).get("sm_mhz") or 1965.0) * 1e6 / 1e12
    r = {"bound": "int", "achieved": achieved, "peak": peak, "unit": "Tiop/s", "frac": achieved / peak,
         # DRAM bytes per launch: ncu --set full on the final round-2 build (profiles/r02_*_ncu_sections.txt): 2.117 GB read + 2.181 GB
         # written per 2^27-point launch of the C2 kernel = 32.0 B/point = the algorithmic scratch write + read (comp / uncomp / eth
         # the same; xpoint with 10^6 targets 37.7 B/point: 3.6 MB bloom + 20 MB table + prefix bitmap on top)
         "traffic": w.get("dram_b_per_point", 32.0) * pts_per_launch,
         "traffic_unit": "bytes of DRAM read+write per launch (ncu-measured %.1f B/point x points per launch)" % w.get("dram_b_per_point", 32.0),
         "kernel": "kh_scan_kernel", "ops_per_point": w["ops"], "launch_ms": launch_ms,
         # `achieved` / `frac` count only the work the kernel still DOES (SURVEY's constant minus what the prefix bitmap, the SHA-256 schedule
         # table and the peeled Keccak rounds made unnecessary); with SURVEY §8d's undiscounted per-point figure the fraction would be:
         "ops_per_point_survey": w["survey_ops"], "frac_survey_ops": achieved * w["survey_ops"] / w["ops"] / peak,
         "peak_source": "measured live: kh_int_peak LOP3+IMAD dual-pipe rate; ALU pipe alone %.2f, IMAD %.2f, IMAD.WIDE %.2f Tiop/s"
                        % (peaks["lop3"] / 1e12, peaks["imad"] / 1e12, peaks["imad_wide"] / 1e12),
         "frac_of_nominal_64_lanes": achieved / nominal, "nominal_peak": nominal,
         "hbm_gbs_scratch": 32.0 * value_pts_s / 1e9, "hbm_peak_gbs": env.mp.get("hbm_gbs")}
    # the pipe that actually binds (ncu, profiles/): ALU ops per point against the live ALU-only rate for the hash kernels,
    # wide multiplies per point against the live IMAD.WIDE.U32.X rate for the x-only walk
    if "alu_ops" in w:
        a = pts_per_launch * w["alu_ops"] / (launch_ms * 1e-3)
        r["binding_pipe"] = {"pipe": "alu", "ops_per_point": w["alu_ops"], "achieved": a / 1e12, "peak": peaks["lop3"] / 1e12,
                             "frac": a / peaks["lop3"], "unit": "Tiop/s"}
    elif "wide_mults" in w:
        a = pts_per_launch * w["wide_mults"] / (launch_ms * 1e-3)
        r["binding_pipe"] = {"pipe": "fma_heavy (IMAD.WIDE.U32.X)", "ops_per_point": w["wide_mults"], "achieved": a / 1e12,
                             "peak": peaks["imad_wide"] / 1e12, "frac": a / peaks["imad_wide"], "unit": "Tiop/s",
                             "note": "wide multiply-adds in carry chains; they do not overlap with ALU-pipe work (kh_pipe_peak: IMAD.WIDE + LOP3 "
                                     "issued together run at the SUM of their times), so the ~400 ALU ops per point of the walk are not free"}
    return r


def cpu_baseline_scan(env, wl, records, flat, seconds):
    """the unmodified reference on the head of the same range (contains planted key index 0) -> also a hit-parity check"""
    w = WORKLOADS[wl]
    try:
        cores = os.cpu_count() or 1
        n_cpu = ref_sample_points(wl, cores, seconds)
        r = run_reference(wl, records, w["start"], n_cpu, cores)
        mine = sorted(f[1] for f in flat if w["start"] <= f[1] < w["start"] + n_cpu) if wl != "c1" else None
        # the reference tests its range cursor outside the mutex (keyhunt.cpp:3314), so racing threads may scan a few
        # chunks past the end: compare inside the sample only
        ref_keys = [k for k in r["keys"] if w["start"] <= k < w["start"] + n_cpu]
        rate = r["steady"] or (n_cpu / r["wall_s"])
        return {"value": rate / 1e6, "unit": "Mkeys/s", "cores": cores, "kind": "reference",
                "sample": "first %d keys of the %s range, %s %s, %s, %.1f s wall; rate = %s" %
                          (n_cpu, wl, os.path.basename(ref_binary()), r["flags"], cpu_model(), r["wall_s"],
                           "its own statistics lines in steady state" if r["steady"] else "sample / wall (too short for steady state)"),
                "wall_value": n_cpu / r["wall_s"] / 1e6,
                "hits_equal_gpu": (mine == ref_keys) if mine is not None else None, "hits": len(ref_keys)}
    except Exception as e:  # the bench line must still come out
        return {"value": None, "unit": "Mkeys/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}


def run_scan_workload(env, wl, Ksteps, W, e2e_steps, cpu_seconds):
    """one scan workload, weak-scaled over the ranks: rank r owns steps [r*K, (r+1)*K) of a contiguous range (no collective on
    the data path).  Returns the result dict on rank 0, None elsewhere."""
    K, kh, torch = env.K, env.kh, env.torch
    w = WORKLOADS[wl]
    world, rank = env.world, env.rank
    mode, crypto, search = kh_modes(K, wl)
    total_points = world * Ksteps * STEP_POINTS
    records, planted = make_targets(kh, K, wl, seed=2, n_points_total=total_points)
    rec_host = torch.frombuffer(bytearray(records), dtype=torch.uint8).pin_memory()     # pinned host copy of the targets
    rec_c = (C.c_char * len(records)).from_address(rec_host.data_ptr())
    shard_start = w["start"] + rank * Ksteps * STEP_POINTS
    kh.set_targets(mode, records, crypto=crypto, search=search)

    # ---- warm-up (sub-ranges outside the timed range; their hits are discarded) --------------------------------
    for s in range(W):
        kh.scan(w["start"] - (s + 1) * STEP_POINTS if w["start"] > (W + 1) * STEP_POINTS else w["start"] + (total_points + s * STEP_POINTS), STEP_POINTS)
    kh.poll_hits()
    kh.stats(reset=True)

    # ---- timed: device-resident ----------------------------------------------------------------------
    clocks = ClockSampler(env.local)
    env.barrier()
    clocks.start()
    t0 = time.perf_counter()
    for s in range(Ksteps):
        kh.scan(shard_start + s * STEP_POINTS, STEP_POINTS)
    hits = kh.poll_hits()
    env.barrier()
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    st = kh.stats(reset=True)
    dev_ms = st["walk_ms"] + st["setup_ms"] + st["aux_ms"]

    # ---- timed: end to end through the C ABI with host buffers ---------------------------------------
    e2e_steps = max(1, min(e2e_steps, Ksteps))
    env.barrier()
    t0 = time.perf_counter()
    e2e_hits = []
    for s in range(e2e_steps):
        kh._ck(kh._lib.kh_set_targets(kh._h, mode, crypto, search, rec_c, len(records) // 20, None, None))   # H2D from pinned memory
        kh.scan(shard_start + s * STEP_POINTS, STEP_POINTS)
        e2e_hits += kh.poll_hits()                                                                           # D2H
    env.barrier()
    e2e_wall = time.perf_counter() - t0
    st_e2e = kh.stats(reset=True)

    # ---- reduce over ranks (max time; hits gathered to rank 0) ---------------------------------------
    found = sorted((h.index, h.key, h.matched.hex(), h.kind) for h in hits)   # index is relative to its step's start
    dev_ms_max, wall_ms_max, e2e_ms_max = env.reduce([dev_ms, wall * 1e3, e2e_wall * 1e3], "max")
    launches = int(env.reduce([st["walk_launches"] + st["other_launches"]], "sum")[0])
    all_found = env.gather(found)
    if rank != 0:
        return None

    flat = [f for part in all_found for f in part]
    got_keys = sorted(f[1] for f in flat)
    if wl == "c1":
        # the reference's own fixture (tests/1to32.txt, puzzle keys 1..32, all below 2^32): every target must be matched exactly
        # once, and the 24 keys the unmodified reference reported for -r 1:FFFFFF (tests/golden/scans.json) must be among them
        want_keys = []
        try:
            gold = json.load(open(os.path.join(ROOT, "tests", "golden", "scans.json")))
            want_keys = sorted(int(k, 16) for c in gold if c["name"] == "address_compress_1to32" for k in c["keys"])
        except Exception:
            pass
        targets = {records[i:i + 20].hex() for i in range(0, len(records), 20)}
        ok = (len(want_keys) == 24 and set(want_keys) <= set(got_keys) and sorted(f[2] for f in flat) == sorted(targets)) if total_points >= (1 << 32) else None
    else:
        want_keys = sorted(k for (_, k) in planted.values())
        ok = (got_keys == want_keys)
    if ok is False:
        log("[bench] %s HIT MISMATCH: got %d keys" % (wl, len(got_keys)))

    cpu = cpu_baseline_scan(env, wl, records, flat, cpu_seconds) if (cpu_seconds and world == 1) else None
    points = world * Ksteps * STEP_POINTS
    value = points / (dev_ms_max * 1e-3) / 1e6
    pts_per_launch = Ksteps * STEP_POINTS / max(1, st["walk_launches"])
    launch_ms = st["walk_ms"] / max(1, st["walk_launches"])
    return {
        "metric": "Mkeys/s (%s, points/s)" % wl, "value": value, "unit": "Mkeys/s", "steps": Ksteps, "warmup": W,
        "ms_per_step": dev_ms_max / Ksteps,
        "config": {"workload": w["desc"], "keys_per_step_per_gpu": STEP_POINTS, "keys_total": points, "targets": len(records) // 20,
                   "l2_note": "inputs larger than L2: every step walks 2^32 new keys; per-thread scratch (9.9 GB) streams through HBM",
                   "displayed_keys_multiplier": w["disp"], "walker_threads": st["walker_threads"], "gpu": env.info["name"]},
        "clocks": clk,
        "e2e": {"value": world * e2e_steps * STEP_POINTS / (e2e_ms_max * 1e-3) / 1e6, "unit": "Mkeys/s",
                "h2d_bytes_per_step": len(records) + 64, "d2h_bytes_per_step": 8 + 160 * max(1, len(e2e_hits)) // max(1, e2e_steps),
                "steps": e2e_steps, "launches": st_e2e["walk_launches"] + st_e2e["other_launches"]},
        "gpu_launches": launches,
        "displayed_keys_value": value * w["disp"],   # the reference multiplies by 2 for -l compress (keyhunt.cpp:2889-2891)
        "wall_ms_per_step": wall_ms_max / Ksteps,
        "hits": {"found": len(got_keys), "all_planted_found_and_nothing_else": ok, "collapsed_batches": st.get("collapsed_batches", 0)},
        "roofline": roofline(env, wl, value * 1e6, pts_per_launch, launch_ms, clk),
        "cpu_baseline": cpu,
    }


def run_c4(env, k, steps, W, cpu_k, cpu_seconds):
    """C4 (BASELINE.json configs[3]): -m bsgs -k 512, n = 2^44, one public key in [2^64, 2^65); a step = one sweep of 2^15
    windows (2^28 giant steps, 2^60 keys)"""
    kh = env.kh
    t0 = time.perf_counter()
    kh.bsgs_build(N44, k)
    build_wall = time.perf_counter() - t0
    st = kh.stats(reset=True)
    d = kh.bsgs_describe()
    env.c4_build_s = build_wall
    build = {"wall_s": build_wall, "baby_walk_ms": st["walk_ms"], "sort_ms": st["aux_ms"], "baby_steps_per_s": d.m / (st["walk_ms"] * 1e-3),
             "tier1_GB": d.tier[0].bytes * 256 / 1e9, "m": d.m, "m2": d.m2, "m3": d.m3, "launches": st["walk_launches"] + st["other_launches"]}
    win, Wn, start = 2 * N44, 1 << 15, 1 << 64
    rnd = random.Random(4)
    key = start + 37 * (1 << 45) + rnd.randrange(1 << 45)      # planted key, found in window 37
    pub = kh.derive([key])[0]
    t0 = time.perf_counter()
    got = kh.bsgs_search((pub.pub_x, pub.pub_y), start, 1 << 65)
    t_find = time.perf_counter() - t0
    for s in range(W):
        kh.bsgs_search((GX, GY), start + s * Wn * win, start + (s + 1) * Wn * win)
    kh.stats(reset=True)
    clocks = ClockSampler(env.local)
    clocks.start()
    t0 = time.perf_counter()
    for s in range(steps):                                       # a key outside the range: every step sweeps all its windows
        a = start + (W + s) * Wn * win
        kh.bsgs_search((GX, GY), a, a + Wn * win)
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    st = kh.stats(reset=True)
    dev_ms = st["walk_ms"] + st["setup_ms"] + st["aux_ms"]
    gs_kernel = st["points"] / (st["walk_ms"] * 1e-3)
    gs_api = st["points"] / (dev_ms * 1e-3)
    hbm = env.mp.get("hbm_gbs", 6650.0)
    cpu = None
    if cpu_seconds:
        try:
            r = reference_bsgs_rate(cpu_k, cpu_seconds)
            if r:
                cpu = {"value": r["giant_steps_per_s"] / 1e6, "unit": "M giant steps/s", "cores": r["cores"], "kind": "reference",
                       "sample": "keyhunt -m bsgs -k %d (m=2^%d) -t %d, %d s of its own statistics lines in steady state after a %.0f s table build, %s; "
                                 "giant steps/s = keys/s / 2m does not depend on k (its k=512 table build alone takes >= 14 min)"
                                 % (r["k"], r["m"].bit_length() - 1, r["cores"], r["sample_s"], r["build_s"] or -1, cpu_model()),
                       "keys_per_s_at_its_k": r["keys_per_s"]}
        except Exception as e:
            cpu = {"value": None, "unit": "M giant steps/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % e}
    return {
        "metric": "M giant steps/s (c4 bsgs -k %d)" % k, "value": gs_api / 1e6, "unit": "M giant steps/s", "steps": steps, "warmup": W,
        "ms_per_step": dev_ms / steps, "keys_per_s": gs_api * 2 * d.m, "kernel_only_giant_steps_per_s": gs_kernel,
        "config": {"workload": "C4 bsgs -k %d, n=2^44 (m=2^%d), 1 public key, range [2^64, 2^65), step = 2^15 windows = 2^28 giant steps"
                               % (k, d.m.bit_length() - 1), "gpu": env.info["name"],
                   "l2_note": "tier-1 bloom (%.1f GB) and its prefix bitmap are far larger than L2; every probe is a random HBM sector" % build["tier1_GB"]},
        "clocks": clk, "build": build, "wall_ms_per_step": wall * 1e3 / steps,
        "planted": {"found": got == key, "time_to_find_s": t_find},
        "e2e": {"value": st["points"] / wall / 1e6, "unit": "M giant steps/s", "h2d_bytes_per_step": 128, "d2h_bytes_per_step": 48, "steps": steps},
        "gpu_launches": st["walk_launches"] + st["other_launches"], "tier1_positives": st["tier1_positives"],
        "roofline": {"bound": "hbm", "achieved": gs_kernel * 64 / 1e9, "peak": hbm, "unit": "GB/s", "frac": gs_kernel * 64 / 1e9 / hbm,
                     "traffic": 97.5 * (st["points"] / max(1, st["walk_launches"])),
                     "kernel": "kh_giant_kernel", "bytes_per_giant_step": 64,
                     "note": "algorithmic 64 B/step (SURVEY §8d: 2 random 32-B sectors; here 16 B + 16 B of prefix-product scratch and one 32-B sector of "
                             "the baby-point prefix bitmap that answers for the tier-1 bloom); ncu: 81 B read + 16 B written per step with the 64-byte L2 "
                             "fetch (145 B with the default 128-byte sector promotion, the SAME kernel time: profiles/r02_giant_probe_flavours.txt) — the "
                             "kernel is bound by the EC arithmetic like the xpoint walk, not by this traffic",
                     "binding_pipe": ({"pipe": "fma_heavy (IMAD.WIDE.U32.X)", "ops_per_point": 224, "achieved": gs_kernel * 224 / 1e12,
                                       "peak": env.peaks["imad_wide"] / 1e12, "frac": gs_kernel * 224 / env.peaks["imad_wide"], "unit": "Tiop/s",
                                       "ops_note": "224 wide multiply-adds of the walk per giant step (ncu executes 255 incl. index arithmetic)"}
                                      if env.peaks else None)},
        "cpu_baseline": cpu,
    }


def strong_c5(env, log2_total):
    """strong scaling of C5: a FIXED 2^log2_total-key range from 0x10000000000 cut into contiguous shards of whole 2^32-key
    chunks (sharding.shard_range); clock (wall, max over ranks) around target upload + set-up + scan + hit gather over NCCL."""
    from keyhunt_b200 import sharding
    K, kh = env.K, env.kh
    total = 1 << log2_total
    out = {}
    for wl in ("c5btc", "c5eth"):
        w = WORKLOADS[wl]
        mode, crypto, search = kh_modes(K, wl)
        records, planted = make_targets(kh, K, wl, seed=5, n_points_total=total)
        my_start, my_n = sharding.shard_range(w["start"], total, env.world, env.rank, chunk=STEP_POINTS)
        kh.stats(reset=True)
        env.barrier()
        t0 = time.perf_counter()
        kh.set_targets(mode, records, crypto=crypto, search=search)
        hits = []
        for off in range(0, my_n, STEP_POINTS):
            kh.scan(my_start + off, STEP_POINTS)
            hits += [(h.key, h.matched.hex()) for h in kh.poll_hits()]
        merged = sharding.gather_hits(env.dist, hits)            # the only collective: a few hit records
        env.barrier()
        wall = time.perf_counter() - t0
        st = kh.stats(reset=True)
        scan_ms = st["walk_ms"] + st["setup_ms"] + st["aux_ms"]
        wall_max, scan_max = env.reduce([wall, scan_ms * 1e-3], "max")
        rate_sum = env.reduce([my_n / (scan_ms * 1e-3) if my_n else 0.0], "sum")[0]
        want = sorted(k for (_, k) in planted.values())
        t1_est = total / (rate_sum / env.world)                  # what ONE of these GPUs needs for the whole range at its device-timed rate
        out[wl] = {"keys": total, "time_s": wall_max, "scan_s_max_rank": scan_max, "value": total / wall_max / 1e6, "unit": "Mkeys/s",
                   "planted_found_and_nothing_else": sorted(k for k, _ in merged) == want, "hits": len(merged),
                   "t1_estimate_s": t1_est, "efficiency": t1_est / (env.world * wall_max),
                   "overhead_s": wall_max - scan_max}
    out["note"] = ("time_s = wall clock (max over ranks) around kh_set_targets + every kh_scan of the rank's shard + the NCCL hit gather; "
                   "efficiency = (range / mean device-timed single-GPU rate) / (N x time_s); what is lost is the per-call set-up (606,208 start "
                   "points per 2^32-key call), the target upload and the gather — there is no collective on the data path")
    return out


def strong_c4(env, k, log2_width):
    """strong scaling of C4: the 2N-key windows of [2^64, 2^64 + 2^log2_width) dealt over the ranks in contiguous blocks
    (sharding.shard_windows); every GPU builds its own tables (reported), clock around the searches + the result gather."""
    from keyhunt_b200 import sharding
    kh = env.kh
    start, width = 1 << 64, 1 << log2_width
    win = 2 * N44
    n_windows = width // win
    env.barrier()
    t0 = time.perf_counter()
    built = False
    try:
        built = kh.bsgs_describe().m == (1 << 22) * k            # the `workloads` block of an N=1 run has built them already
    except Exception:
        pass
    if not built:
        kh.bsgs_build(N44, k)
    env.barrier()
    build_s = env.c4_build_s if built else time.perf_counter() - t0
    d = kh.bsgs_describe()
    kh.stats(reset=True)
    first, count = sharding.shard_windows(n_windows, env.world, env.rank)
    rnd = random.Random(9)
    key = start + rnd.randrange(width)
    pub = kh.derive([key])[0]
    res = {}
    for name, target in (("sweep", (GX, GY)), ("planted", (pub.pub_x, pub.pub_y))):
        kh.stats(reset=True)
        env.barrier()
        t0 = time.perf_counter()
        got = kh.bsgs_search(target, start + first * win, start + (first + count) * win) if count else None
        found = sharding.gather_hits(env.dist, [got] if got is not None else [])
        env.barrier()
        wall = time.perf_counter() - t0
        st = kh.stats(reset=True)
        dev_s = (st["walk_ms"] + st["setup_ms"] + st["aux_ms"]) * 1e-3
        wall_max, dev_max, setup_max = env.reduce([wall, dev_s, st["setup_ms"] * 1e-3], "max")
        steps_sum, walk_rate_sum = env.reduce([st["points"], st["points"] / max(1e-9, st["walk_ms"] * 1e-3)], "sum")
        res[name] = {"time_s": wall_max, "device_s_max_rank": dev_max, "setup_s_max_rank": setup_max, "giant_steps": steps_sum,
                     "value": steps_sum / wall_max / 1e6, "unit": "M giant steps/s", "found": found}
        if name == "sweep":
            t1 = steps_sum / (walk_rate_sum / env.world)
            res[name]["t1_estimate_s"] = t1
            res[name]["efficiency"] = t1 / (env.world * wall_max)
        else:
            res[name]["planted_found"] = (found == [key])
    return {"k": k, "range_log2": log2_width, "windows": n_windows, "build_s_every_gpu": build_s, "m": d.m, **res,
            "note": "windows dealt in contiguous blocks, tables built on every GPU; sweep = a key outside the range (every window walked); "
                    "efficiency = (giant steps / mean kernel rate of one GPU) / (N x time_s): the loss is the start-point set-up of every "
                    "kh_bsgs_search call (one scalar multiplication per walker) and launch tails, not communication"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-workloads", action="store_true", help="skip the `workloads` block (the other BASELINE configs)")
    ap.add_argument("--no-strong", action="store_true", help="skip the `strong` block (fixed-range C5 / C4 over the ranks)")
    ap.add_argument("--strong-log2", type=int, default=37, help="keys of the fixed C5 range = 2^this per currency")
    ap.add_argument("--bsgs-k", type=int, default=512)
    ap.add_argument("--steps-per-launch", type=int, default=0, help="library option (profiling runs use 1 for short kernels)")
    ap.add_argument("--step-points-log2", type=int, default=32, help="keys per step = 2^this (profiling runs use less)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # not under torchrun: relaunch one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.warmup < 3:
        log("[bench] warm-up raised to 3 (timing rules)")
        args.warmup = 3
    global STEP_POINTS
    STEP_POINTS = 1 << args.step_points_log2

    env = Env(args)
    rank = env.rank
    env.peaks = env.kh.int_peak()
    try:
        env.pipe = env.kh.pipe_peak()
    except Exception as e:
        log("[bench] kh_pipe_peak failed: %s" % e)
    t_start = time.perf_counter()

    # ---- the headline workload --------------------------------------------------------------------------
    line = run_scan_workload(env, args.workload, args.steps, args.warmup, args.e2e_steps, 0 if args.no_cpu_baseline else 12)
    if rank == 0:
        line = {**{"metric": line["metric"], "value": line["value"], "unit": line["unit"], "n_gpus": world, "steps": line["steps"],
                   "warmup": line["warmup"], "ms_per_step": line["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                   "vs_baseline": None, "dtype": "u32", "data": "synthetic"}, **{k: v for k, v in line.items() if k not in ("metric", "value", "unit", "steps", "warmup", "ms_per_step")}}
        line["peaks"] = {"unit": "thread-ops/s, whole chip, measured live by kh_int_peak / kh_pipe_peak", **env.peaks, **(env.pipe or {})}

    # ---- the other BASELINE configs (N = 1) --------------------------------------------------------------
    if world == 1 and not args.no_side_workloads:
        side = {}
        for wl in ("c1", "c3", "c5btc", "c5eth"):
            if wl == args.workload:
                continue
            try:
                t0 = time.perf_counter()
                r = run_scan_workload(env, wl, 3, 3, 2, 0 if args.no_cpu_baseline else 6.5)
                r["block_wall_s"] = time.perf_counter() - t0
                side[wl] = r
                log("[bench] %s: %.0f Mkeys/s, e2e %.0f, hits ok %s" % (wl, r["value"], r["e2e"]["value"], r["hits"]["all_planted_found_and_nothing_else"]))
            except Exception as e:
                side[wl] = {"error": str(e)}
        try:
            t0 = time.perf_counter()
            r = run_c4(env, args.bsgs_k, 8, 3, 16, 0 if args.no_cpu_baseline else 12)
            r["block_wall_s"] = time.perf_counter() - t0
            side["c4"] = r
            log("[bench] c4: %.0f M giant steps/s (kernel %.0f), planted found %s" % (r["value"], r["kernel_only_giant_steps_per_s"] / 1e6, r["planted"]["found"]))
        except Exception as e:
            side["c4"] = {"error": str(e)}
        line["workloads"] = side

    # ---- strong scaling: fixed ranges over the ranks -------------------------------------------------------
    if not args.no_strong:
        strong = {}
        try:
            strong["c5"] = strong_c5(env, args.strong_log2)
        except Exception as e:
            strong["c5"] = {"error": str(e)}
        try:
            strong["c4"] = strong_c4(env, args.bsgs_k, 66)
        except Exception as e:
            strong["c4"] = {"error": str(e)}
        if rank == 0:
            strong["scaling"] = "strong"
            strong["n_gpus"] = world
            line["strong"] = strong

    if rank == 0:
        line["bench_wall_s"] = time.perf_counter() - t_start
        print(json.dumps(line))
    env.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
