"""GPU x >= 2 (skipped on a one-GPU box): the N > 1 paths on hardware — the CLI's `-t N` (= N GPUs) for scan modes and for BSGS
(windows dealt in contiguous blocks, tables built on every GPU), and two library contexts on two devices.  The one-GPU run of
the same command is the comparison; the golden files say what the unmodified reference found."""
import json
import os
import shutil
import subprocess
import tempfile

import pytest

import keyhunt_b200 as K

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CLI = os.path.join(ROOT, "keyhunt_b200", "keyhunt-b200")


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")


def _run(args, cwd):
    r = subprocess.run([CLI] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    return r.returncode, r.stdout


def _records(cwd, per):
    fn = os.path.join(cwd, "KEYFOUNDKEYFOUND.txt")
    if not os.path.exists(fn):
        return []
    lines = open(fn).read().splitlines()
    return sorted("|".join(lines[i:i + per]) for i in range(0, len(lines), per))


@pytest.fixture()
def dirs():
    a, b = tempfile.mkdtemp(prefix="khm_1_"), tempfile.mkdtemp(prefix="khm_2_")
    yield a, b
    shutil.rmtree(a, ignore_errors=True)
    shutil.rmtree(b, ignore_errors=True)


@needs2
def test_cli_bsgs_on_two_gpus(dirs):
    one, two = dirs
    gold = json.load(open(os.path.join(GOLD, "bsgs.json")))
    for d in (one, two):
        open(os.path.join(d, "p.txt"), "w").write("\n".join(gold["pubkeys"]) + "\n")
    args = ["-m", "bsgs", "-f", "p.txt", "-n", "0x400000", "-k", "2", "-r", "100000:10000000000", "-q"]
    rc1, out1 = _run(args + ["-t", "1"], one)
    rc2, out2 = _run(args + ["-t", "2"], two)
    assert "All points were found" in out1 and "All points were found" in out2, out2[-2000:]
    assert "2 x " in out2                                  # the driver really opened two devices
    assert _records(one, 2) == _records(two, 2) and len(_records(two, 2)) == len(gold["pubkeys"])
    keys = sorted(int(r.split("|")[0].split()[-1], 16) for r in _records(two, 2))
    assert keys == sorted(int(k, 16) for k in gold["keys"])        # what the unmodified reference found


@needs2
def test_cli_scan_on_two_gpus(dirs):
    one, two = dirs
    args = ["-m", "address", "-f", GOLD + "/1to32.txt", "-r", "1:FFFFFFFFF", "-l", "compress", "-n", "0x10000000", "-q"]
    rc1, out1 = _run(args + ["-t", "1"], one)
    rc2, out2 = _run(args + ["-t", "2"], two)
    assert rc1 == 0 and rc2 == 0 and "End" in out2, out2[-2000:]
    assert _records(one, 4) == _records(two, 4) and len(_records(two, 4)) == 32     # puzzle keys 1..32 are all below 2^32


@needs2
def test_two_contexts_split_a_bsgs_range(oracle):
    from keyhunt_b200 import sharding
    n, k = 1 << 24, 2
    start, n_windows = 0x8000000001, 64
    win = 2 * n
    keys = [start + 5 * win + 123, start + 40 * win + 9, start + 63 * win + win - 1]
    with K.KeyHunt(0) as a, K.KeyHunt(1) as b:
        for kh in (a, b):
            kh.bsgs_build(n, k)
        for key in keys:
            pub = oracle.pubkey(key)
            got = []
            for rank, kh in enumerate((a, b)):
                first, count = sharding.shard_windows(n_windows, 2, rank)
                r = kh.bsgs_search(pub, start + first * win, start + (first + count) * win)
                if r is not None:
                    got.append((rank, r))
            assert [g[1] for g in got] == [key]
            assert got[0][0] == (0 if (key - start) // win < 32 else 1)
