"""GPU: the BSGS server (keyhunt_b200/keyhunt-b200-bsgsd) against the UNMODIFIED reference
server (oracle/_ref/bsgsd, CPU) started side by side with the same -n/-k: every request gets byte-identical replies
(line protocol and HTTP POST), and each server starts from the table files the other one wrote."""
import json
import os
import re
import shutil
import signal
import socket
import subprocess
import tempfile
import time

import pytest

from _oracle import Oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_D = os.path.join(ROOT, "keyhunt_b200", "keyhunt-b200-bsgsd")
REF_D = os.path.join(ROOT, "oracle", "_ref", "bsgsd")
NK = ["-n", "0x1000000", "-k", "4"]          # m = 16384 baby steps, windows of 2^25 keys


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class Server:
    def __init__(self, exe, cwd, extra=()):
        self.port = free_port()
        self.log = open(os.path.join(cwd, "server_%d.log" % self.port), "w+")
        self.proc = subprocess.Popen([exe] + NK + ["-p", str(self.port), "-i", "127.0.0.1", "-t", "1"] + list(extra), cwd=cwd,
                                     stdout=self.log, stderr=subprocess.STDOUT)
        deadline = time.time() + 300
        while time.time() < deadline:                  # ready = the port accepts (the reference's stdout is block-buffered into the log)
            if self.proc.poll() is not None:
                raise RuntimeError("server exited: " + self.text()[-2000:])
            try:
                socket.create_connection(("127.0.0.1", self.port), timeout=1).close()
                return
            except OSError:
                time.sleep(0.2)
        tail = self.text()[-2000:]
        self.stop()
        raise RuntimeError("server did not start: " + tail)

    def text(self):
        self.log.flush()
        self.log.seek(0)
        return self.log.read()

    def ask(self, payload):
        s = socket.create_connection(("127.0.0.1", self.port), timeout=120)
        s.sendall(payload)
        data = b""
        while True:
            try:
                d = s.recv(4096)
            except ConnectionResetError:
                break
            if not d:
                break
            data += d
        s.close()
        return data

    def stop(self):
        if self.proc.poll() is None:
            self.proc.send_signal(signal.SIGTERM)      # the exact PID we started
            try:
                self.proc.wait(timeout=20)
            except subprocess.TimeoutExpired:
                self.proc.kill()
                self.proc.wait()
        self.log.close()


def post(body):
    return b"POST / HTTP/1.1\r\nHost: x\r\nContent-Type: application/json\r\nContent-Length: %d\r\n\r\n" % len(body) + body


def mask(reply):
    return re.sub(rb"X-Elapsed-Seconds: [0-9.]+", b"X-Elapsed-Seconds: T", reply)


@pytest.fixture(scope="module")
def keys():
    o = Oracle()
    out = {}
    for k in (0x1234567, 0x1FFFFFF, 0x2000001, 0x2000000, 0x123456789, 1):
        x, y = o.pubkey(k)
        xb, yb = x.to_bytes(32, "big"), y.to_bytes(32, "big")
        out[k] = ((b"03" if y & 1 else b"02") + xb.hex().encode(), b"04" + (xb + yb).hex().encode())
    return out


def requests(keys):
    c, u = keys[0x1234567]
    edge, _ = keys[0x1FFFFFF]
    past, _ = keys[0x2000001]
    base, _ = keys[0x2000000]
    far, _ = keys[0x123456789]
    one, _ = keys[1]
    J = lambda p, a, b: post(b'{"pubkey":"%s","from":"%s","to":"%s"}' % (p, a, b))
    return [
        c + b" 1000000:2000000\n",                    # found, compressed key
        c + b" 2000000:3000000\n",                    # not in range
        u + b" 1000000 2000000\n",                    # uncompressed key, three-token form
        c.upper() + b" 1000000:2000000\r\n",          # upper-case hex, CRLF
        c + b" 1000000\n",                            # no range end -> 400
        c + b" xx:yy\n",                              # not hex -> 400
        b"\n",                                        # nothing -> 400
        c + b" 0:2000000\n",                          # start 0
        c + b" 1234567:1234568\n",                    # one-key range: a whole window is searched
        c + b" 1234568:1334568\n",                    # key just below the start
        c + b" 2000000:1000000\n",                    # start > end -> no window
        c + b" 1000000:1000000\n",                    # empty
        edge + b" 1:2\n",                             # window overshoot past `to`: key 0x1FFFFFF is inside window [1, 1+2^25)
        past + b" 1:2\n",                             # first key after that window
        past + b" 1:2000002\n",                       # ... which a second window reaches
        base + b" 2000000:2000001\n",                 # key == window base
        far + b" 100000000:200000000\n",              # 2^28-key range = 8 windows, found in the second
        far + b" 1:100000000\n",                      # 128 windows, not found
        one + b" 1:1000000\n",                        # key 1 at range start 1
        one + b" 0:1000000\n",
        J(c, b"1000000", b"2000000"),
        J(c, b"2000000", b"3000000"),
        J(u, b"1", b"FFFFFFFF"),
        post(b"{}"),                                  # missing fields -> HTTP 400
        post(b'{"pubkey":"%s","from":"zz","to":"10"}' % c),
    ]


def test_replies_match_golden_reference_replies(keys):
    """tests/golden/bsgsd_replies.json: what the unmodified reference server answered to the same requests"""
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "bsgsd_replies.json")))
    assert [gold["n"], gold["k"]] == [NK[1], NK[3]]
    rq = requests(keys)
    assert [c["request"].encode("latin1") for c in gold["cases"]] == rq
    d = tempfile.mkdtemp(prefix="bsgsd_gpu_")
    g = None
    try:
        g = Server(GPU_D, d)
        for q, c in zip(rq, gold["cases"]):
            assert mask(g.ask(q)) == c["reply"].encode("latin1"), q
    finally:
        if g:
            g.stop()
        shutil.rmtree(d, ignore_errors=True)


@pytest.mark.skipif(not os.path.exists(REF_D), reason="oracle/_ref/bsgsd not built")
def test_replies_identical_to_reference_server(keys):
    gdir, rdir = tempfile.mkdtemp(prefix="bsgsd_gpu_"), tempfile.mkdtemp(prefix="bsgsd_ref_")
    g = r = None
    try:
        g = Server(GPU_D, gdir)
        r = Server(REF_D, rdir, ["-6"])
        for q in requests(keys):
            a, b = g.ask(q), r.ask(q)
            assert mask(a) == mask(b), (q, a, b)
        assert g.ask(keys[0x1234567][0] + b" 1000000:2000000\n") == b"1234567\n"
        assert b"200 OK" in g.ask(post(b'{"pubkey":"%s","from":"1","to":"2000000"}' % keys[0x1234567][0]))
        assert "Key found privkey 1234567" in g.text()
    finally:
        for s in (g, r):
            if s:
                s.stop()
        shutil.rmtree(gdir, ignore_errors=True)
        shutil.rmtree(rdir, ignore_errors=True)


@pytest.mark.skipif(not os.path.exists(REF_D), reason="oracle/_ref/bsgsd not built")
def test_servers_start_from_each_others_table_files(keys):
    """bsgsd always works from keyhunt_bsgs_*.blm/.tbl: ours loads the files the reference wrote (checksums verified)
    and the reference loads ours; the files are identical up to the heap pointer the reference leaves in each header."""
    gdir, rdir = tempfile.mkdtemp(prefix="bsgsd_gpu_"), tempfile.mkdtemp(prefix="bsgsd_ref_")
    names = ["keyhunt_bsgs_4_16384.blm", "keyhunt_bsgs_6_512.blm", "keyhunt_bsgs_7_16.blm", "keyhunt_bsgs_2_16.tbl"]
    q = keys[0x1234567][0] + b" 1000000:2000000\n"
    try:
        g = Server(GPU_D, gdir, ["-S"]); g.stop()    # -S: writes the four files
        r = Server(REF_D, rdir, ["-6"]); r.stop()
        for n in names:
            a, b = open(os.path.join(gdir, n), "rb").read(), open(os.path.join(rdir, n), "rb").read()
            assert len(a) == len(b), n
            if n.endswith(".blm"):
                rec = len(a) // 256
                for sh in range(256):
                    ra, rb = bytearray(a[sh * rec:(sh + 1) * rec]), bytearray(b[sh * rec:(sh + 1) * rec])
                    ra[64:72] = rb[64:72] = b"\0" * 8      # struct bloom.bf is a heap pointer of the writing process
                    assert ra == rb, (n, sh)
            else:
                body = lambda x: sorted(x[i:i + 16] for i in range(0, len(x) - 32, 16))   # the reference sort is not stable
                assert body(a) == body(b), n
        g2 = Server(GPU_D, rdir, ["-S"])             # our server on the reference's files, checksums on
        try:
            assert "Reading bP Table from file" in g2.text() and "Writing" not in g2.text()
            assert g2.ask(q) == b"1234567\n"
        finally:
            g2.stop()
        r2 = Server(REF_D, gdir)                     # the reference on our files, checksums on (no -6)
        try:
            assert "Reading bP Table from file" in r2.text()
            assert r2.ask(q) == b"1234567\n"
        finally:
            r2.stop()
    finally:
        shutil.rmtree(gdir, ignore_errors=True)
        shutil.rmtree(rdir, ignore_errors=True)


def test_bad_public_key_is_refused_not_fatal(keys):
    """the reference's ParsePublicKeyHex exits the whole server on a malformed key; ours answers 400 and keeps serving"""
    d = tempfile.mkdtemp(prefix="bsgsd_gpu_")
    g = None
    try:
        g = Server(GPU_D, d)
        assert g.ask(b"zz 1:2\n") == b"400 Bad Request"
        assert g.ask(b"02" + b"00" * 32 + b" 1:2\n") == b"400 Bad Request"          # x = 0 is not on the curve
        assert g.ask(post(b'{"pubkey":"05aa","from":"1","to":"2"}')).startswith(b"HTTP/1.1 400 Bad Request")
        assert g.ask(b"x" * 5000 + b"\n") == b"400 Bad Request"                       # over-long line
        assert g.ask(keys[1][0] + b" 1:1000\n") == b"1\n"
    finally:
        if g:
            g.stop()
        shutil.rmtree(d, ignore_errors=True)
