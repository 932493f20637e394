"""CPU: the host drivers parse the reference's command line before they touch the GPU, and without a GPU they fail loudly —
there is no CPU fallback in the product.  (Skipped on a box that has a GPU: there the same command lines are run for real by
tests/test_gpu_cli.py, test_gpu_vanity.py and test_gpu_bsgsd.py.)"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "keyhunt_b200", "keyhunt-b200")
BSGSD = os.path.join(ROOT, "keyhunt_b200", "keyhunt-b200-bsgsd")


def _has_gpu():
    if not shutil.which("nvidia-smi"):
        return False
    try:
        return subprocess.run(["nvidia-smi", "-L"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=20).stdout.strip() != b""
    except Exception:
        return False


pytestmark = pytest.mark.skipif(_has_gpu() or not os.path.exists(CLI), reason="needs the built CLI and a box without a GPU")


def run(exe, args):
    return subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60, cwd="/tmp")


def test_scan_modes_refuse_to_run_without_a_gpu():
    for args in (["-m", "rmd160", "-f", "x.rmd", "-r", "1:FFFF", "-l", "both"], ["-m", "xpoint", "-f", "x.txt", "-b", "40"],
                 ["-m", "bsgs", "-f", "p.txt", "-r", "1:FFFFFFFF", "-k", "2", "-n", "0x400000"]):
        r = run(CLI, args)
        assert r.returncode != 0 and "no usable CUDA device (this build has no CPU fallback)" in r.stdout, r.stdout


def test_reference_option_surface_is_parsed_first():
    r = run(CLI, ["-m", "vanity", "-v", "1Bit", "-v", "0notbase58", "-v", "1" * 30, "-l", "compress", "-r", "1:100000", "-e", "-q"])
    for line in ("[+] Mode vanity", "[+] Added Vanity search : 1Bit", 'The string "0notbase58" is not Valid Base58',
                 '[+] Vanity search "%s" was NOT Added' % ("1" * 30), "[+] Search compress only", "[+] Endomorphism enabled", "[+] Quiet thread output"):
        assert line in r.stdout, (line, r.stdout)
    r = run(CLI, ["-m", "bsgs", "-B", "backward", "-f", "p.txt", "-r", "1:FFFFFFFF", "-k", "2", "-n", "0x400000"])
    assert "[+] Mode BSGS backward" in r.stdout
    r = run(CLI, ["-m", "bsgs", "-B", "dance", "-f", "p.txt", "-r", "1:FFFFFFFF"])
    assert r.returncode != 0 and "-B dance is not supported" in r.stdout
    r = run(CLI, ["-m", "address", "-f", "a.txt", "-r", "1:FFFF", "-k", "4", "-n", "0x100000"])      # validate_nk (util.c:358)
    assert r.returncode != 0 and "k value 4 is too large for n 0x100000 (max 1)" in r.stdout
    r = run(CLI, ["-m", "minikeys", "-f", "a.txt"])
    assert r.returncode != 0 and "not part of the GPU path" in r.stdout
    r = run(CLI, ["-m", "address", "-f", "a.txt", "-R"])
    assert r.returncode != 0 and "-R (random mode)" in r.stdout


def test_bsgsd_refuses_to_start_without_a_gpu():
    r = run(BSGSD, ["-k", "2", "-n", "0x400000", "-p", "18099"])
    assert r.returncode != 0 and "no usable CUDA device" in r.stdout and "Listening" not in r.stdout
