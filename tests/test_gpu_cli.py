"""GPU: the drop-in command-line driver (keyhunt_b200/keyhunt-b200) against the UNMODIFIED reference binary
(oracle/_ref/keyhunt, CPU) run side by side on the same command lines: identical KEYFOUNDKEYFOUND.txt records
and identical `-S` BSGS table files."""
import hashlib
import json
import os
import shutil
import subprocess
import tempfile

import pytest

from _oracle import REF_BIN

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CLI = os.path.join(ROOT, "keyhunt_b200", "keyhunt-b200")
SCANS = {c["name"]: c for c in json.load(open(os.path.join(GOLD, "scans.json")))}


def run(exe, args, cwd):
    r = subprocess.run([exe] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    return r.returncode, r.stdout


def records(cwd, per):
    """KEYFOUNDKEYFOUND.txt as a sorted list of `per`-line records (hit order varies with threads)"""
    fn = os.path.join(cwd, "KEYFOUNDKEYFOUND.txt")
    if not os.path.exists(fn):
        return []
    lines = open(fn).read().splitlines()
    return sorted("|".join(lines[i:i + per]) for i in range(0, len(lines), per))


@pytest.fixture()
def dirs():
    a, b = tempfile.mkdtemp(prefix="khcli_gpu_"), tempfile.mkdtemp(prefix="khcli_ref_")
    yield a, b
    shutil.rmtree(a, ignore_errors=True)
    shutil.rmtree(b, ignore_errors=True)


CASES = [
    ("address_compress", ["-m", "address", "-f", GOLD + "/1to32.txt", "-r", "1:FFFFFF", "-l", "compress", "-n", "0x100000"], 4),
    ("rmd160_both", ["-m", "rmd160", "-f", GOLD + "/1to32.rmd", "-r", "1:FFFFFF", "-l", "both", "-n", "0x100000"], 4),   # ranges end where a racy extra chunk of the reference (cursor checked outside its mutex) cannot contain a target
    ("address_eth", ["-m", "address", "-c", "eth", "-f", GOLD + "/1to32.eth", "-r", "1:FFFFFF", "-n", "0x100000"], 2),
    ("xpoint", ["-m", "xpoint", "-f", GOLD + "/substracted40.txt", "-r", "8000000000:8003000000", "-n", "0x100000"], 4),
    ("stride", ["-m", "rmd160", "-f", GOLD + "/1to32.rmd", "-r", "1:FFFFF", "-l", "compress", "-n", "0x100000", "-I", "3"], 4),
]


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/keyhunt not built")
@pytest.mark.parametrize("name,args,per", CASES)
def test_cli_output_identical_to_reference(dirs, name, args, per):
    g, r = dirs
    rc_g, out_g = run(CLI, args + ["-q", "-t", "1"], g)
    rc_r, out_r = run(REF_BIN, args + ["-q", "-s", "0", "-t", str(os.cpu_count() or 4)], r)
    assert rc_g == 0 and "End" in out_g, out_g[-2000:]
    assert rc_r == 0 and "End" in out_r, out_r[-2000:]
    rec_g, rec_r = records(g, per), records(r, per)
    assert rec_g == rec_r
    assert len(rec_g) >= 1


def test_cli_planted_uncompressed_and_opposite_parity(dirs):
    g, _ = dirs
    for name, flag in (("planted_uncompress", "uncompress"), ("planted_opposite_parity", "compress")):
        fn = os.path.join(g, name + ".rmd")
        open(fn, "w").write("\n".join(SCANS[name]["targets"]) + "\n")
        if os.path.exists(os.path.join(g, "KEYFOUNDKEYFOUND.txt")):
            os.remove(os.path.join(g, "KEYFOUNDKEYFOUND.txt"))
        rc, out = run(CLI, ["-m", "rmd160", "-f", fn, "-r", "2000000000000000:2000000000400000", "-l", flag, "-n", "0x100000", "-q"], g)
        assert rc == 0, out[-2000:]
        keys = sorted(int(rec.split("|")[0].split(":")[1], 16) for rec in records(g, 4))
        assert keys == sorted(int(k, 16) for k in SCANS[name]["keys"])


def test_cli_c1_full_sweep_checksum(dirs):
    """BASELINE config 1 end to end: the canonicalised KEYFOUNDKEYFOUND.txt of the full -r 1:FFFFFFFF sweep has the
    SHA-256 the reference run produced (cb75e1c2...c6ffbb)"""
    g, _ = dirs
    rc, out = run(CLI, ["-m", "address", "-f", GOLD + "/1to32.txt", "-r", "1:FFFFFFFF", "-l", "compress", "-q"], g)
    assert rc == 0 and "End" in out, out[-2000:]
    recs = records(g, 4)
    assert len(recs) == 32
    digest = hashlib.sha256(("\n".join(recs) + "\n").encode()).hexdigest()
    assert digest == "cb75e1c290cfb3c669936eebbcdfb7ef1f192ca578edcbcd03e8371471c6ffbb"


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/keyhunt not built")
def test_cli_bsgs_files_and_keys_identical_to_reference(dirs):
    g, r = dirs
    pubs = [l.strip() for l in open(os.path.join(GOLD, "bsgs_pubkeys.txt")) if l.strip()]
    for d in (g, r):
        open(os.path.join(d, "p.txt"), "w").write("\n".join(pubs) + "\n")
    args = ["-m", "bsgs", "-f", "p.txt", "-n", "0x400000", "-k", "2", "-r", "100000:10000000000", "-S", "-q"]
    rc_g, out_g = run(CLI, args + ["-t", "1"], g)
    rc_r, out_r = run(REF_BIN, args + ["-s", "0", "-t", str(os.cpu_count() or 4)], r)
    assert "All points were found" in out_g, out_g[-2000:]
    assert "All points were found" in out_r, out_r[-2000:]
    assert rc_g == rc_r == 1                      # sic: success exits with EXIT_FAILURE (keyhunt.cpp:4855-4858)
    assert records(g, 2) == records(r, 2) and len(records(g, 2)) == len(pubs)
    files = sorted(f for f in os.listdir(r) if f.startswith("keyhunt_bsgs_"))
    assert files == sorted(f for f in os.listdir(g) if f.startswith("keyhunt_bsgs_")) and len(files) == 4
    for fn in files:
        a, b = open(os.path.join(g, fn), "rb").read(), open(os.path.join(r, fn), "rb").read()
        assert len(a) == len(b), fn
        if fn.endswith(".blm"):
            rec = len(a) // 256
            for s in range(256):
                ra, rb = bytearray(a[s * rec:(s + 1) * rec]), bytearray(b[s * rec:(s + 1) * rec])
                ra[64:72] = rb[64:72] = b"\0" * 8          # struct bloom.bf is a heap pointer of the writing process
                assert ra == rb, (fn, s)
        else:
            body = lambda x: sorted(x[i:i + 16] for i in range(0, len(x) - 32, 16))   # reference sort is not stable (ties)
            assert body(a) == body(b)
            assert a[-32:] == hashlib.sha256(a[:-32]).digest()
    # second run loads the files instead of rebuilding and still finds the keys
    os.remove(os.path.join(g, "KEYFOUNDKEYFOUND.txt"))
    rc, out = run(CLI, args + ["-t", "1"], g)
    assert "tables loaded from files" in out and "All points were found" in out


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/keyhunt not built")
def test_cli_target_cache_file_identical_and_interchangeable(dirs):
    """-S in scan modes: data_<sha256 prefix>.dat (bloom + sorted table, keyhunt.cpp:7756) is byte-identical to the
    reference's, and each tool loads the other's file and reports the same keys"""
    g, r = dirs
    args = ["-m", "rmd160", "-f", GOLD + "/1to32.rmd", "-r", "1:FFFFF", "-l", "compress", "-n", "0x100000", "-S", "-q"]
    rc_g, out_g = run(CLI, args + ["-t", "1"], g)
    rc_r, out_r = run(REF_BIN, args + ["-s", "0", "-t", "1"], r)   # -t 1: the reference checks the cursor outside its mutex, extra threads overshoot
    assert rc_g == 0 and rc_r == 0, (out_g[-1000:], out_r[-1000:])
    fg = [f for f in os.listdir(g) if f.startswith("data_")]
    fr = [f for f in os.listdir(r) if f.startswith("data_")]
    assert fg == fr and len(fg) == 1
    a, b = bytearray(open(os.path.join(g, fg[0]), "rb").read()), bytearray(open(os.path.join(r, fr[0]), "rb").read())
    assert len(a) == len(b)
    a[32 + 64:32 + 72] = b[32 + 64:32 + 72] = b"\0" * 8      # struct bloom.bf: heap pointer of the writer
    assert a == b
    want = records(r, 4)
    assert records(g, 4) == want and len(want) >= 10
    # cross-load: swap the cache files, run again with -S: both must say they read the file and find the same keys
    shutil.copy(os.path.join(r, fr[0]), os.path.join(g, fg[0]))
    open(os.path.join(r, fr[0]), "wb").write(bytes(a))       # ours (pointer field zero) into the reference directory
    for d in (g, r):
        os.remove(os.path.join(d, "KEYFOUNDKEYFOUND.txt"))
    rc_g, out_g = run(CLI, args + ["-t", "1"], g)
    rc_r, out_r = run(REF_BIN, args + ["-s", "0", "-t", "1"], r)   # -t 1: the reference checks the cursor outside its mutex, extra threads overshoot
    assert "Reading file data_" in out_g and "Reading file data_" in out_r
    assert records(g, 4) == want and records(r, 4) == want


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/keyhunt not built")
@pytest.mark.parametrize("mode", ["sequential", "backward", "both"])
def test_cli_bsgs_window_pickers_identical_to_reference(dirs, oracle, mode):
    """-B sequential | backward | both (keyhunt.cpp:4549, :5953, :6211): same found keys as the reference, including the
    keys that only one alignment of the 2N windows can see (just past the range end, at the range start)"""
    g, r = dirs
    inside = [0x1000002, 0x1234567, 0x17FFFFF, 0x1800000, 0x2800065, 0x8000000, 0x8FFFFFF]
    edge = [0x1000001, 0x9000000, 0x9000001, 0x9000010, 0xFFFFFF]          # alignment-dependent or never found
    keys = inside + ([] if mode == "both" else edge)                         # `both` picks its windows with rand() in the reference
    lines = []
    for k in keys:
        x, y = oracle.pubkey(k)
        lines.append(("03" if y & 1 else "02") + "%064x" % x)
    for d in (g, r):
        open(os.path.join(d, "p.txt"), "w").write("\n".join(lines) + "\n")
    args = ["-m", "bsgs", "-f", "p.txt", "-n", "0x400000", "-k", "2", "-r", "1000001:9000000", "-q", "-B", mode]
    rc_g, out_g = run(CLI, args + ["-t", "1"], g)
    rc_r, out_r = run(REF_BIN, args + ["-s", "0", "-t", "2"], r)
    assert rc_g == rc_r, (out_g[-1500:], out_r[-1500:])
    assert records(g, 2) == records(r, 2), (records(g, 2), records(r, 2))
    found = {int(rec.split("|")[0].split()[-1], 16) for rec in records(g, 2)}
    assert set(inside) <= found


def test_cli_small_n_keeps_the_gpu_full(dirs, kh):
    """`-n 0x1000000` (the chunk size SURVEY §8d prescribes for the CPU run) used to give the GPU 16,384 of
    606,208 walkers per kh_scan.  Contiguous chunk claims are now scanned as one run: the rate must be within 3 % of the
    default `-n 0x100000000`, and the records identical (C2-like: rmd160 -l both, 1,024 targets, 12 planted, 2^34 keys)."""
    import random
    import re
    g, r = dirs
    start, n = 0x2000000000000000, 1 << 34
    rnd = random.Random(77)
    idxs = sorted({0, n - 1} | {rnd.randrange(n) for _ in range(10)})
    infos = kh.derive([start + i for i in idxs])
    recs = [(inf.h160_uncomp if j % 2 else inf.h160_comp) for j, inf in enumerate(infos)] + [rnd.randbytes(20) for _ in range(1012)]
    rates, found = {}, {}
    for d, nflag in ((g, "0x1000000"), (r, "0x100000000")):
        fn = os.path.join(d, "t.rmd")
        open(fn, "w").write("".join(x.hex() + "\n" for x in recs))
        rc, out = run(CLI, ["-m", "rmd160", "-f", fn, "-r", "%x:%x" % (start, start + n), "-l", "both", "-n", nflag, "-q", "-t", "1"], d)
        assert rc == 0 and "End" in out, out[-2000:]
        m = re.search(r"Total (\d+) keys in \d+ seconds: .*\((\d+) keys/s\)", out)
        assert m and int(m.group(1)) == n, out[-500:]
        rates[nflag], found[nflag] = int(m.group(2)), records(d, 4)
    print("\nCLI rmd160 -l both over 2^34 keys: -n 0x1000000 -> %.0f Mkeys/s, -n 0x100000000 -> %.0f Mkeys/s"
          % (rates["0x1000000"] / 1e6, rates["0x100000000"] / 1e6))
    try:        # evidence for profiles/ (the GPU box merges gpurun_out/ back)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump({"what": "keyhunt-b200 -m rmd160 -l both, 1024 targets, 2^34 keys, keys/s printed by the CLI itself (context + target upload included)",
                   "rate_n_0x1000000": rates["0x1000000"], "rate_n_0x100000000": rates["0x100000000"], "records_identical": found["0x1000000"] == found["0x100000000"]},
                  open(os.path.join(ROOT, "gpurun_out", "cli_small_n.json"), "w"))
    except OSError:
        pass
    assert found["0x1000000"] == found["0x100000000"] and len(found["0x1000000"]) == len(idxs)
    assert rates["0x1000000"] >= 0.97 * rates["0x100000000"], rates
