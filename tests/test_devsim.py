"""CPU: the product's device headers (keyhunt_b200/csrc/*.cuh) compiled for the host (tests/devsim) agree
with the oracle: limb arithmetic, message packers, hashes, bloom index arithmetic, batch geometry with
interleaved walker threads and multi-launch continuation, and the fused scan/emit logic."""
import ctypes as C
import os
import random
import subprocess

import pytest

from _oracle import (CRYPTO_BTC, CRYPTO_ETH, MODE_ADDRESS, MODE_RMD160, MODE_XPOINT, N_ORDER, P_FIELD, SEARCH_BOTH,
                     SEARCH_COMPRESS, SEARCH_UNCOMPRESS, be32)

HERE = os.path.dirname(os.path.abspath(__file__))


class DsHit(C.Structure):
    _fields_ = [("index", C.c_uint64), ("kind", C.c_uint32), ("matched", C.c_uint8 * 20), ("variant", C.c_uint32)]


@pytest.fixture(scope="module")
def ds():
    d = os.path.join(HERE, "devsim")
    subprocess.check_call(["make", "-C", d], stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(d, "libkh_devsim.so"))
    lib.ds_xxh64_20.restype = C.c_uint64; lib.ds_xxh64_20.argtypes = [C.c_char_p, C.c_uint64]
    lib.ds_xxh64_32.restype = C.c_uint64; lib.ds_xxh64_32.argtypes = [C.c_char_p, C.c_uint64]
    lib.ds_bloom_mod.restype = C.c_uint64; lib.ds_bloom_mod.argtypes = [C.c_uint64, C.c_uint64]
    lib.ds_scan.restype = C.c_int64
    lib.ds_scan.argtypes = [C.c_int, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_uint32, C.c_char_p, C.c_char_p,
                            C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(DsHit), C.c_uint32, C.c_int]
    lib.ds_walk_dump.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_char_p]
    return lib


def _f2(fn, a, b):
    o = C.create_string_buffer(32); fn(be32(a), be32(b), o); return int.from_bytes(o.raw, "big")


def _f1(fn, a):
    o = C.create_string_buffer(32); fn(be32(a), o); return int.from_bytes(o.raw, "big")


def test_field_limb_algorithms(ds):
    rnd = random.Random(7)
    P = P_FIELD
    edge = [0, 1, 2, P - 1, P - 2, 2**255, (1 << 224) - 1, 2**32, 2**32 + 977, P - 977, 0xFFFFFFFF, P - 2**32, 2**256 - 2**33]
    vals = [e % P for e in edge] + [rnd.randrange(P) for _ in range(400)]
    for i, a in enumerate(vals):
        b = vals[(i * 7 + 3) % len(vals)]
        assert _f2(ds.ds_fe_mul, a, b) == a * b % P
        assert _f2(ds.ds_fe_add, a, b) == (a + b) % P
        assert _f2(ds.ds_fe_sub, a, b) == (a - b) % P
        assert _f1(ds.ds_fe_sqr, a) == a * a % P
    for a in vals[:24]:
        assert _f1(ds.ds_fe_inv, a) == pow(a, P - 2, P)


def test_scalar_mult_and_hashes(ds, oracle):
    rnd = random.Random(8)
    for k in [1, 2, 3, 0xDEADBEEF, N_ORDER - 1] + [rnd.randrange(1, N_ORDER) for _ in range(6)]:
        o = C.create_string_buffer(64); ds.ds_pubkey(be32(k), o)
        assert (int.from_bytes(o.raw[:32], "big"), int.from_bytes(o.raw[32:], "big")) == oracle.pubkey(k)
    out = C.create_string_buffer(20)
    for _ in range(200):
        x, y = rnd.randrange(P_FIELD), rnd.randrange(P_FIELD)
        for pre in (2, 3):
            ds.ds_hash160_comp(pre, be32(x), out); assert out.raw == oracle.hash160_comp(pre, x)
        ds.ds_hash160_uncomp(be32(x) + be32(y), out); assert out.raw == oracle.hash160_uncomp(x, y)
        ds.ds_hash160_uncomp_tab(be32(x) + be32(y), out); assert out.raw == oracle.hash160_uncomp(x, y)
        ds.ds_eth_addr(be32(x) + be32(y), out); assert out.raw == oracle.eth_addr(x, y)
        d, s = rnd.randbytes(32), rnd.randrange(2**64)
        assert ds.ds_xxh64_20(d[:20], s) == oracle.xxh64(d[:20], s)
        assert ds.ds_xxh64_32(d, s) == oracle.xxh64(d, s)


def test_sha256_schedule_table_of_the_uncompressed_second_block(ds, oracle):
    """hash.cuh KH_SHA_UNC2_TAB: the second SHA-256 block of 04||X||Y holds one data byte; its 256 precomputed message schedules
    give the same hash160 as computing the schedule, for every value of that byte"""
    out = C.create_string_buffer(20)
    rnd = random.Random(10)
    x = rnd.randrange(P_FIELD)
    for v in range(256):
        y = (rnd.randrange(P_FIELD) & ~0xFF) | v
        ds.ds_hash160_uncomp_tab(be32(x) + be32(y), out)
        assert out.raw == oracle.hash160_uncomp(x, y), v


def test_bloom_exact_modulo(ds):
    rnd = random.Random(9)
    for _ in range(20000):
        bits = rnd.choice([287551, 28755175, 241215892, 7537996, 2**20, 2**33 + 5, rnd.randrange(2, 2**40), 3, 2**63 + 11, 2**64 - 1])
        x = rnd.choice([rnd.randrange(2**64), 2**64 - 1, (bits * 5) % 2**64, (bits * 5 - 1) % 2**64, 0, bits - 1, bits])
        assert ds.ds_bloom_mod(x, bits) == x % bits


@pytest.mark.parametrize("start,stride,nb,T,spl", [(1, 1, 5, 3, 2), (0x8000000000, 1, 4, 4, 1), (0xABCDEF0123456789ABCDEF, 0x1F3, 3, 2, 5),
                                                     (12345, 7, 7, 3, 1)])
def test_walk_geometry(ds, oracle, start, stride, nb, T, spl):
    """interleaved walker threads (batch t, t+T, ...), several launches, ragged tail == the reference's batches"""
    out = C.create_string_buffer(nb * 1024 * 64)
    ds.ds_walk_dump(be32(start), be32(stride), nb, T, spl, out)
    for b in range(nb):
        assert out.raw[b * 65536:(b + 1) * 65536] == oracle.batch_points(start + b * 1024 * stride, stride, True)


def test_two_level_centre_setup(ds, oracle):
    """setup.cuh: walker 64*i + j starts on A_i + B_j (63 affine additions behind one inversion per row) — the same points as one
    scalar multiplication per walker, for scans, strides, giant walks (negated step, base point Q), ragged last rows, and rows
    where a difference is zero (A_i = +-B_j: the whole row falls back) or a walker sits at infinity (reported)."""
    ds.ds_setup_centres.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_int, C.c_char_p, C.POINTER(C.c_uint64)]
    ninf = C.c_uint64()
    q = oracle.pubkey(0xC0FFEE1234567)
    qxy = be32(q[0]) + be32(q[1])
    for k0, s, T, neg, base in [(1, 1, 130, 0, None), (0x2000000000000000, 1, 256, 0, None), (0xDEADBEEF, 977, 65, 0, None),
                                (0x8000000000 + (1 << 31), 1 << 32, 200, 1, qxy), (5, 3, 64, 1, qxy), (7, 1, 1, 0, None), (7, 1, 63, 0, None)]:
        assert ds.ds_setup_centres(be32(k0), be32(s), T, neg, base, C.byref(ninf)) == 0, (k0, s, T, neg)
        assert ninf.value == 0
    # centre of walker 64 = 5120*G = B_5: row 1 has a zero difference (doubling case) and must fall back; on the way there
    # walker 59 crosses key n (infinity, reported by row 0)
    assert ds.ds_setup_centres(be32(N_ORDER - 60928), be32(1), 130, 0, None, C.byref(ninf)) == 0 and ninf.value == 1
    # the same coincidence without any walker at infinity: stride 3, walker 64 = 5120*3*G
    assert ds.ds_setup_centres(be32(N_ORDER - 60928 * 3 + 1), be32(3), 130, 0, None, C.byref(ninf)) == 0 and ninf.value == 0
    assert ds.ds_setup_centres(be32(N_ORDER - 60928 * 3), be32(3), 130, 0, None, C.byref(ninf)) == 0 and ninf.value == 1
    # centre of walker 64 = -(3*1024)*G = -B_3: walker 67 is the point at infinity, everything else is right
    assert ds.ds_setup_centres(be32(N_ORDER - 66048 - 3072), be32(1), 130, 0, None, C.byref(ninf)) == 0 and ninf.value == 1
    # the row BASE at infinity (walker 64 = key 0): reported by phase 1, the row's other walkers are still right where defined
    assert ds.ds_setup_centres(be32(N_ORDER - 66048), be32(1), 130, 0, None, C.byref(ninf)) == 0 and ninf.value == 1


def _last_plan(ds):
    out = (C.c_uint64 * 3)()
    ds.ds_last_plan(out)
    return dict(segments=out[0], collapsed=out[1], distinct_T=out[2])


def test_plan_constant(ds):
    assert ds.ds_plan_selfcheck() == 1


def test_walk_centre_on_the_hop_point(ds, oracle):
    """start = 512, stride 1, T walkers: batch T-1 has its centre ON W = T*1024*G, so the hop's difference would be zero
    (ADVICE r1).  The reference has no such hop: its batch is fine, and so must ours be: the host plan (plan.hpp) cuts the scan
    there and walks the rest with another T, the kernels never see the coincidence."""
    T, nb = 4, 12
    out = C.create_string_buffer(nb * 65536)
    ds.ds_walk_dump(be32(512), be32(1), nb, T, 2, out)
    plan = _last_plan(ds)
    assert plan["collapsed"] == 0 and plan["segments"] == 2 and plan["distinct_T"] == 2
    for b in range(1, nb):
        assert out.raw[b * 65536:(b + 1) * 65536] == oracle.batch_points(512 + b * 1024, 1, True), b
    # batch 0 has its centre on 1024*G = the reference's own next-start delta (_2Gn, keyhunt.cpp:3448): the REFERENCE's batch
    # collapses there; ours holds the true points (a superset of what the reference can find)
    for i in (0, 1, 511, 512, 513, 1023):
        x, y = oracle.pubkey(512 + i)
        assert out.raw[64 * i:64 * i + 64] == be32(x) + be32(y)
    assert out.raw[:65536] != oracle.batch_points(512, 1, True)
    # the same with a stride: start + 512*stride = T*1024*stride  <=>  start = (T*1024 - 512) * stride - 1024*b*stride
    stride = 977
    ds.ds_walk_dump(be32(512 * stride), be32(stride), nb, T, 3, out)
    assert _last_plan(ds)["segments"] == 2
    for b in range(1, nb):
        assert out.raw[b * 65536:(b + 1) * 65536] == oracle.batch_points(512 * stride + b * 1024 * stride, stride, True), b


def test_walk_batch_without_inverse_matches_the_reference(ds, oracle):
    """a range that runs over key 0 (mod n): the centre of a batch is +-512*G, one difference is zero, no shared inverse
    exists.  IntGroup::ModInv then yields zeros and the reference's 1023 non-centre points are deterministic garbage
    (SURVEY App. B.11); the same garbage comes out here (the batch is planned as a segment of its own so that its walker's
    useless next centre is never used), it is counted, and every other batch is right."""
    start, T, nb = N_ORDER - 1024, 2, 6
    out = C.create_string_buffer(nb * 65536)
    ds.ds_walk_dump(be32(start), be32(1), nb, T, 3, out)
    plan = _last_plan(ds)
    assert plan["collapsed"] == 2 and plan["segments"] == 3
    for b in range(nb):
        base = start + b * 1024
        assert out.raw[b * 65536:(b + 1) * 65536] == oracle.batch_points(base % N_ORDER if b else base, 1, True), b


def test_plan_is_a_single_segment_for_ordinary_ranges(ds):
    out = C.create_string_buffer(3 * 65536)
    for start, stride in ((1, 1), (0x2000000000000000, 1), (0xDEADBEEF12345, 977), (2**255 + 12345, 2**64 + 1)):
        ds.ds_walk_dump(be32(start), be32(stride), 3, 2, 2, out)
        assert _last_plan(ds) == dict(segments=1, collapsed=0, distinct_T=1), (start, stride)


KINDS = {"xpoint": (0, MODE_XPOINT, CRYPTO_BTC, SEARCH_COMPRESS), "comp": (1, MODE_RMD160, CRYPTO_BTC, SEARCH_COMPRESS),
         "uncomp": (2, MODE_RMD160, CRYPTO_BTC, SEARCH_UNCOMPRESS), "both": (3, MODE_ADDRESS, CRYPTO_BTC, SEARCH_BOTH),
         "eth": (4, MODE_ADDRESS, CRYPTO_ETH, SEARCH_COMPRESS)}


@pytest.mark.parametrize("name", sorted(KINDS))
def test_scan_emit_logic(ds, oracle, name):
    kind, mode, crypto, search = KINDS[name]
    rnd = random.Random(hash(name) & 0xFFFF)
    start, stride, nb, T = 0x2000000000000777, 3, 6, 4
    n = nb * 1024
    recs = []
    for j, i in enumerate(sorted({0, 511, 512, 513, 1023, 1024, n - 1} | {rnd.randrange(n) for _ in range(10)})):
        x, y = oracle.pubkey(start + i * stride)
        if name == "xpoint":
            recs.append(be32(x)[:20])
        elif name == "eth":
            recs.append(oracle.eth_addr(x, y))
        elif name == "comp":
            recs.append(oracle.hash160_comp((2 + (y & 1)) if j % 3 else (3 - (y & 1)), x))
        elif name == "uncomp":
            recs.append(oracle.hash160_uncomp(x, y))
        else:
            recs.append(oracle.hash160_uncomp(x, y) if j % 2 else oracle.hash160_comp(2 + (y & 1), x))
    recs += [rnd.randbytes(20) for _ in range(100)]
    t = oracle.targets_new(b"".join(recs))
    want = oracle.scan(t, mode, crypto, search, start, stride, n, nthreads=2)
    bl = oracle.targets_bloom(t)
    d = oracle.bloom_desc(bl)
    hits = (DsHit * 256)()
    cnt = ds.ds_scan(kind, oracle.targets_table(t), len(recs), oracle.bloom_bytes(bl), d["bits"], d["hashes"], be32(start),
                     be32(stride), nb, T, 2, hits, 256, 0)
    oracle.targets_free(t)
    got = sorted((hits[i].index, hits[i].kind, bytes(hits[i].matched)) for i in range(cnt))
    assert got == sorted((h["index"], h["kind"], h["matched"]) for h in want)
    assert cnt >= 10


@pytest.mark.parametrize("name", sorted(KINDS))
def test_scan_emit_logic_endomorphism(ds, oracle, name):
    """-e candidates (x, beta*x, beta^2*x; +-y; the ETH slot-4 quirk): raw device hits (index, kind, l, matched) == oracle"""
    from _oracle import BETA, BETA2
    kind, mode, crypto, search = KINDS[name]
    rnd = random.Random(hash(name) & 0xFFF)
    start, stride, nb, T = 0x3000000000000321, 1, 3, 2
    n = nb * 1024
    P = P_FIELD
    recs = []
    for j, i in enumerate(sorted({0, 512, 1023, 1024, n - 1} | {rnd.randrange(n) for _ in range(13)})):
        x, y = oracle.pubkey(start + i * stride)
        xv = [x, x * BETA % P, x * BETA2 % P][j % 3]
        yy = y if (j // 3) % 2 == 0 else P - y
        if name == "xpoint":
            recs.append(be32(xv)[:20])
        elif name == "eth":
            recs.append(oracle.eth_addr(xv, yy))
        elif name == "comp" or (name == "both" and j % 2):
            recs.append(oracle.hash160_comp(2 + (yy & 1), xv))
        else:
            recs.append(oracle.hash160_uncomp(xv, yy))
    recs += [rnd.randbytes(20) for _ in range(60)]
    t = oracle.targets_new(b"".join(recs))
    want = oracle.scan(t, mode, crypto, search, start, stride, n, nthreads=2, endo=True)
    bl = oracle.targets_bloom(t)
    d = oracle.bloom_desc(bl)
    hits = (DsHit * 256)()
    cnt = ds.ds_scan(kind, oracle.targets_table(t), len(recs), oracle.bloom_bytes(bl), d["bits"], d["hashes"], be32(start),
                     be32(stride), nb, T, 2, hits, 256, 1)
    oracle.targets_free(t)
    got = sorted((hits[i].index, hits[i].kind, hits[i].variant, bytes(hits[i].matched)) for i in range(cnt))
    assert got == sorted((h["index"], h["kind"], h["variant"], h["matched"]) for h in want)
    assert cnt >= 10


def _vanity_words(A, B):
    """what kh_set_vanity uploads (kh_scan.cu): 16-bit prefix bitmap + limits as big-endian words"""
    n = len(A) // 20
    van = [0] * (2048 + 10 * n)
    for i in range(n):
        a, b = A[20 * i:20 * i + 20], B[20 * i:20 * i + 20]
        for k in range(5):
            van[2048 + 10 * i + k] = int.from_bytes(a[4 * k:4 * k + 4], "big")
            van[2048 + 10 * i + 5 + k] = int.from_bytes(b[4 * k:4 * k + 4], "big")
        if a <= b:
            for p in range(int.from_bytes(a[:2], "big"), int.from_bytes(b[:2], "big") + 1):
                van[p >> 5] |= 1 << (p & 31)
    return (C.c_uint32 * len(van))(*van), n


@pytest.mark.parametrize("name", ["comp", "uncomp", "both"])
def test_scan_emit_logic_vanity(ds, oracle, name):
    """-m vanity: the device interval test (vanity_match) on the walk's digests == the oracle's vanityrmdmatch, for
    the reference's own prefix decoding (addvanity) of short prefixes that hit within a few thousand keys"""
    kind, mode, crypto, search = KINDS[name]
    A, B, mn, counts = oracle.addvanity(["1A", "1Bi", "1zz", "12"])
    assert min(counts) >= 1
    t = oracle.targets_new_vanity(A, B, mn)
    start, stride, nb, T = 0x5000000000000123, 1, 8, 4
    want = oracle.scan(t, mode, crypto, search, start, stride, nb * 1024, nthreads=2, max_hits=8192)
    oracle.targets_free(t)
    van, n = _vanity_words(A, B)
    ds.ds_set_vanity.argtypes = [C.POINTER(C.c_uint32), C.c_uint32]
    ds.ds_set_vanity.restype = None
    ds.ds_set_vanity(van, n)
    try:
        hits = (DsHit * 8192)()
        dummy = b"\0" * 20
        cnt = ds.ds_scan(kind, dummy, 1, b"\0" * 64, 512, 1, be32(start), be32(stride), nb, T, 2, hits, 8192, 0)
    finally:
        ds.ds_set_vanity(None, 0)
    got = sorted((hits[i].index, hits[i].kind, bytes(hits[i].matched)) for i in range(cnt))
    assert got == sorted((h["index"], h["kind"], h["matched"]) for h in want)
    assert cnt >= 100
