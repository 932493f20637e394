#!/usr/bin/env python3
"""Generates tests/golden/*.json FROM THE REFERENCE ITSELF (run in the build container, where
/root/reference exists and `make -C oracle ref` has produced oracle/_ref/):

  * primitives.json  — white-box vectors out of the reference's own object code (oracle/_ref/libkh_ref.so):
                       ModMulK1 / ModSquareK1 / ModInv, ComputePublicKey, GetHash160_fromX, GetHash160,
                       generate_binaddress_eth, XXH64, bloom_init2 sizing, bloom_add images, the 1024-point
                       batch of thread_process.
  * scans.json       — black-box runs of the unmodified reference binary (oracle/_ref/keyhunt) on the
                       reference's fixture files and on planted synthetic targets: the private keys it reports.
  * bsgs.json        — the reference's `-S` files (3-tier blooms + bP table) as SHA-256 digests, and the keys
                       it finds.

The vectors are committed; the GPU box has no /root/reference, so tests read only these files.
Usage: python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import random
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _oracle import N_ORDER, P_FIELD, REF_BIN, Oracle, RefHarness, be32  # noqa: E402

REFTESTS = "/root/reference/tests"


def hx(b):
    return b.hex()


def primitives(r):
    rnd = random.Random(20261018)
    out = {}
    edge = [0, 1, 2, P_FIELD - 1, P_FIELD - 2, 2**255, (1 << 224) - 1, 2**32 + 977, 0xFFFFFFFF]
    vals = [e % P_FIELD for e in edge] + [rnd.randrange(P_FIELD) for _ in range(40)]
    out["fe_mul"] = [[hex(a), hex(b), hex(r.fe_mul(a, b))] for a, b in zip(vals, vals[3:] + vals[:3])]
    out["fe_sqr"] = [[hex(a), hex(r.fe_sqr(a))] for a in vals]
    out["fe_inv"] = [[hex(a), hex(r.fe_inv(a))] for a in vals[:20]]
    keys = [1, 2, 3, 7, 0xFFFFFFFF, 2**64 + 5, N_ORDER - 1] + [rnd.randrange(1, N_ORDER) for _ in range(20)]
    out["pubkey"] = [[hex(k), hex(r.pubkey(k)[0]), hex(r.pubkey(k)[1])] for k in keys]
    hs = []
    for k in keys:
        x, y = r.pubkey(k)
        hs.append({"key": hex(k), "c02": hx(r.hash160_comp(2, x)), "c03": hx(r.hash160_comp(3, x)),
                   "unc": hx(r.hash160_uncomp(x, y)), "eth": hx(r.eth_addr(x, y)),
                   "scalar_comp": hx(r.hash160_scalar(True, x, y)), "scalar_unc": hx(r.hash160_scalar(False, x, y))})
    out["hashes"] = hs
    xs = []
    for _ in range(40):
        d = rnd.randbytes(32)
        s = rnd.randrange(2**64)
        xs.append([hx(d), hex(s), hex(r.xxh64(d[:20], s)), hex(r.xxh64(d, s))])
    out["xxh64"] = xs
    bl = []
    for e in [1000, 10000, 12345, 33000, 262144, 1000000, 8388608, 33554432]:
        h = r.bloom_new(e)
        bl.append(r.bloom_desc(h))
        r.bloom_free(h)
    out["bloom_sizing"] = bl
    out["sizeof_struct_bloom"] = r._sizeof_bloom()
    imgs = []
    for e, n20, n32, seed in [(10000, 1500, 0, 1), (10000, 0, 1500, 2), (33000, 20000, 0, 3)]:
        rr = random.Random(seed)
        h = r.bloom_new(e)
        items = [rr.randbytes(20) for _ in range(n20)] + [rr.randbytes(32) for _ in range(n32)]
        for it in items:
            r.bloom_add(h, it)
        probes = [rr.randbytes(20 if n20 else 32) for _ in range(3000)]
        fp = [i for i, p in enumerate(probes) if r.bloom_check(h, p)]
        imgs.append({"entries": e, "n20": n20, "n32": n32, "seed": seed, "sha256": hashlib.sha256(r.bloom_bytes(h)).hexdigest(),
                     "false_positive_probe_indices": fp})
        r.bloom_free(h)
    out["bloom_images"] = imgs
    bt = []
    for base, stride in [(1, 1), (0x8000000000, 1), (0xDEADBEEF12345, 977), (rnd.randrange(2**200), rnd.randrange(1, 2**40))]:
        raw = r.batch_points(base, stride, True)
        bt.append({"base": hex(base), "stride": hex(stride), "sha256_xy": hashlib.sha256(raw).hexdigest(),
                   "x0": raw[:32].hex(), "x512": raw[512 * 64:512 * 64 + 32].hex(), "x1023": raw[1023 * 64:1023 * 64 + 32].hex()})
    out["batches"] = bt
    return out


def run_ref(args, cwd):
    r = subprocess.run([REF_BIN] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return r.stdout


def found_keys(cwd):
    fn = os.path.join(cwd, "KEYFOUNDKEYFOUND.txt")
    recs = []
    if os.path.exists(fn):
        lines = open(fn).read().splitlines()
        for i, ln in enumerate(lines):
            if ln.startswith("Private Key:"):
                recs.append(ln.split(":")[1].strip())
            if ln.startswith("Key found privkey"):
                recs.append(ln.split()[3])
        os.remove(fn)
    return sorted(recs, key=lambda s: int(s, 16))


def scans(o):
    out = []
    d = tempfile.mkdtemp(prefix="khgold_")
    try:
        def case(name, args, note=""):
            stdout = run_ref(args + ["-q", "-s", "0", "-t", "8"], d)
            assert "End" in stdout, stdout[-1500:]
            out.append({"name": name, "args": " ".join(a.replace(d + "/", "").replace(REFTESTS + "/", "tests/") for a in args),
                        "keys": found_keys(d), "note": note})
            print(name, len(out[-1]["keys"]))

        # reference fixtures (copied beside this script: 1to32.rmd/.txt/.eth)
        case("rmd160_compress_1to32", ["-m", "rmd160", "-f", REFTESTS + "/1to32.rmd", "-r", "1:FFFFFF", "-l", "compress", "-n", "0x100000"])
        case("address_compress_1to32", ["-m", "address", "-f", REFTESTS + "/1to32.txt", "-r", "1:FFFFFF", "-l", "compress", "-n", "0x100000"])
        case("address_eth_1to32", ["-m", "address", "-c", "eth", "-f", REFTESTS + "/1to32.eth", "-r", "1:FFFFFF", "-n", "0x100000"])
        case("xpoint_substracted40", ["-m", "xpoint", "-f", REFTESTS + "/substracted40.txt", "-r", "8000000000:8010000000", "-n", "0x100000"],
             "README.md:416-437 known answer 800258a2ce (and 8009c16fb9)")
        # planted synthetic targets (files written by this script; the reference decides what is a hit)
        rnd = random.Random(42)
        start, n = 0x2000000000000000, 1 << 22
        idx = [0, 1, 1023, 1024, n - 1] + [rnd.randrange(n) for _ in range(7)]
        unc, mixed, opp, strided = [], [], [], []
        for j, i in enumerate(idx):
            x, y = o.pubkey(start + i)
            unc.append(o.hash160_uncomp(x, y).hex())
            mixed.append((o.hash160_uncomp(x, y) if j % 2 else o.hash160_comp(2 + (y & 1), x)).hex())
            opp.append(o.hash160_comp(3 - (y & 1), x).hex())      # opposite parity: the reference reports n - k
            xs, ys = o.pubkey(start + i * 977)
            strided.append(o.hash160_comp(2 + (ys & 1), xs).hex())
        for nm, lst in [("unc", unc), ("mixed", mixed), ("opp", opp), ("strided", strided)]:
            with open(os.path.join(d, nm + ".rmd"), "w") as f:
                f.write("\n".join(lst + [rnd.randbytes(20).hex() for _ in range(50)]) + "\n")
        rng = "%x:%x" % (start, start + n)
        case("planted_uncompress", ["-m", "rmd160", "-f", d + "/unc.rmd", "-r", rng, "-l", "uncompress", "-n", "0x100000"])
        case("planted_both", ["-m", "rmd160", "-f", d + "/mixed.rmd", "-r", rng, "-l", "both", "-n", "0x100000"])
        case("planted_opposite_parity", ["-m", "rmd160", "-f", d + "/opp.rmd", "-r", rng, "-l", "compress", "-n", "0x100000"])
        for c in out[-3:]:
            c["targets"] = open(os.path.join(d, {"planted_uncompress": "unc", "planted_both": "mixed",
                                                 "planted_opposite_parity": "opp"}[c["name"]] + ".rmd")).read().split()
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return out


def endo_scans(o):
    """-e runs of the unmodified reference on planted endomorphic targets -> scans_endo.json"""
    from _oracle import BETA, BETA2
    P = P_FIELD
    out = []
    d = tempfile.mkdtemp(prefix="khgold_")
    try:
        rnd = random.Random(77)
        start, n = 0x3000000000000000, 1 << 21
        idx = [0, 1, 1023, 1024, n - 1] + [rnd.randrange(n) for _ in range(19)]
        comp, unc, eth, xp = [], [], [], []
        for j, i in enumerate(idx):
            x, y = o.pubkey(start + i)
            xs = [x, x * BETA % P, x * BETA2 % P]
            v = j % 3
            pre = (2 + (y & 1)) if (j // 3) % 2 == 0 else (3 - (y & 1))          # real and opposite parity
            comp.append(o.hash160_comp(pre, xs[v]).hex())
            yy = y if (j // 3) % 2 == 0 else P - y
            unc.append(o.hash160_uncomp(xs[v], yy).hex())
            eth.append("0x" + o.eth_addr(xs[v], yy).hex())
            xp.append("%064x" % xs[v])
        decoys = [rnd.randbytes(20).hex() for _ in range(40)]
        files = {"comp": comp + decoys, "unc": unc + decoys, "both": comp[:12] + unc[12:] + decoys,
                 "eth": eth + ["0x" + t for t in decoys], "xp": xp + [rnd.randbytes(32).hex() for _ in range(40)]}
        for nm, lst in files.items():
            open(os.path.join(d, nm + ".txt"), "w").write("\n".join(lst) + "\n")
        rng = "%x:%x" % (start, start + n)
        cases = [("endo_compress", ["-m", "rmd160", "-f", d + "/comp.txt", "-l", "compress"], "comp"),
                 ("endo_uncompress", ["-m", "rmd160", "-f", d + "/unc.txt", "-l", "uncompress"], "unc"),
                 ("endo_both", ["-m", "rmd160", "-f", d + "/both.txt", "-l", "both"], "both"),
                 ("endo_eth", ["-m", "address", "-c", "eth", "-f", d + "/eth.txt"], "eth"),
                 ("endo_xpoint", ["-m", "xpoint", "-f", d + "/xp.txt"], "xp")]
        for name, args, nm in cases:
            stdout = run_ref(args + ["-e", "-r", rng, "-n", "0x100000", "-q", "-s", "0", "-t", "2"], d)
            assert "End" in stdout, stdout[-1500:]
            fn = os.path.join(d, "KEYFOUNDKEYFOUND.txt")
            lines = open(fn).read().splitlines() if os.path.exists(fn) else []
            per = 2 if nm == "eth" else 4
            recs = sorted("|".join(lines[i:i + per]) for i in range(0, len(lines), per))
            if os.path.exists(fn):
                os.remove(fn)
            out.append({"name": name, "args": " ".join(a.replace(d + "/", "") for a in args) + " -e -r " + rng + " -n 0x100000",
                        "targets": files[nm], "records": recs})
            print(name, len(recs))
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return out


def bsgs():
    d = tempfile.mkdtemp(prefix="khgold_")
    out = {}
    try:
        pubs = [l.strip() for l in open(REFTESTS + "/1to63_65.txt") if l.strip()]
        with open(os.path.join(d, "p.txt"), "w") as f:
            f.write("\n".join(pubs[20:32]) + "\n")
        stdout = run_ref(["-m", "bsgs", "-f", os.path.join(d, "p.txt"), "-n", "0x400000", "-k", "2", "-r", "100000:10000000000",
                          "-t", "8", "-S", "-q", "-s", "0"], d)
        assert "All points were found" in stdout, stdout[-1500:]
        out["args"] = "-m bsgs -f <tests/1to63_65.txt lines 21..32> -n 0x400000 -k 2 -r 100000:10000000000 -S"
        out["pubkeys"] = pubs[20:32]
        out["keys"] = found_keys(d)
        files = {}
        for fn in sorted(os.listdir(d)):
            if fn.startswith("keyhunt_bsgs_"):
                raw = open(os.path.join(d, fn), "rb").read()
                if fn.endswith(".blm"):
                    # 256 x (struct bloom 112 B | bf | 2 x sha256(bf)) — keyhunt.cpp:2519-2530
                    rec = len(raw) // 256
                    nbytes = rec - 112 - 64
                    bf = b"".join(raw[s * rec + 112:s * rec + 112 + nbytes] for s in range(256))
                    hdr = raw[:25]
                    files[fn] = {"shard_bytes": nbytes, "sha256_all_shards": hashlib.sha256(bf).hexdigest(),
                                 "entries": int.from_bytes(hdr[0:8], "little"), "bits": int.from_bytes(hdr[8:16], "little"),
                                 "hashes": hdr[24],
                                 "shard_checksums_ok": all(hashlib.sha256(raw[s * rec + 112:s * rec + 112 + nbytes]).digest() ==
                                                           raw[s * rec + 112 + nbytes:s * rec + 144 + nbytes] for s in range(256))}
                else:
                    body = raw[:-32]
                    ents = sorted((body[i:i + 6].hex(), int.from_bytes(body[i + 8:i + 16], "little")) for i in range(0, len(body), 16))
                    files[fn] = {"entries": ents, "checksum_ok": hashlib.sha256(body).digest() == raw[-32:]}
        out["files"] = files
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return out


if __name__ == "__main__":
    r = RefHarness()
    o = Oracle()
    json.dump(primitives(r), open(os.path.join(HERE, "primitives.json"), "w"), indent=0)
    json.dump(scans(o), open(os.path.join(HERE, "scans.json"), "w"), indent=0)
    json.dump(bsgs(), open(os.path.join(HERE, "bsgs.json"), "w"), indent=0)
    json.dump(endo_scans(o), open(os.path.join(HERE, "scans_endo.json"), "w"), indent=0)
    print("golden vectors written")
