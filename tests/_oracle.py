"""ctypes bindings for the CPU oracle (oracle/_build/libkh_oracle.so, prefix kho_) and, when it has
been built in the container that holds /root/reference, the white-box reference harness
(oracle/_ref/libkh_ref.so, prefix khr_).  TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's
cpu_baseline leg import this module; nothing under keyhunt_b200/ may."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libkh_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libkh_ref.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "keyhunt")
REF_BIN_V3 = os.path.join(ORACLE_DIR, "_ref", "keyhunt_v3")

N_ORDER = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
# -e constants as the reference sets them (keyhunt.cpp:928-931)
LAMBDA = 0x5363ad4cc05c30e0a5261c028812645a122e22ea20816678df02967c1b23bd72
LAMBDA2 = 0xac9c52b33fa3cf1f5ad9e3fd77ed9ba4a880b9fc8ec739c2e0cfc810b51283ce
BETA = 0x7ae96a2b657c07106e64479eac3434e99cf0497512f58995c1396c28719501ee
BETA2 = 0x851695d49a83f8ef919bb86153cbcb16630fb68aed0a766a3ec693d68e6afa40
P_FIELD = 2**256 - 2**32 - 977

MODE_XPOINT, MODE_ADDRESS, MODE_RMD160 = 0, 1, 2
CRYPTO_BTC, CRYPTO_ETH = 0, 1
SEARCH_UNCOMPRESS, SEARCH_COMPRESS, SEARCH_BOTH = 0, 1, 2
HIT_COMP02, HIT_COMP03, HIT_UNCOMP, HIT_ETH, HIT_XPOINT = 0, 1, 2, 3, 4


def be32(v: int) -> bytes:
    return int(v).to_bytes(32, "big")


class Hit(C.Structure):
    _fields_ = [("key_be", C.c_uint8 * 32), ("matched", C.c_uint8 * 20), ("kind", C.c_uint8),
                ("pad", C.c_uint8 * 3), ("index", C.c_uint64)]


class BpEntry(C.Structure):
    _fields_ = [("value", C.c_uint8 * 6), ("pad", C.c_uint8 * 2), ("index", C.c_uint64)]


def build_oracle():
    """(re)build libkh_oracle.so if missing or stale."""
    src = os.path.join(ORACLE_DIR, "kh_oracle.c")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


class _Lib:
    """Common wrapper; `p` is the symbol prefix ('kho_' or 'khr_')."""

    def __init__(self, path, p):
        self.lib = C.CDLL(path)
        self.p = p
        L = self.lib
        u8p = C.c_char_p
        self._sig("fe_mul", None, [u8p, u8p, u8p])
        self._sig("fe_sqr", None, [u8p, u8p])
        self._sig("fe_inv", None, [u8p, u8p])
        self._sig("fe_add", None, [u8p, u8p, u8p])
        self._sig("fe_sub", None, [u8p, u8p, u8p])
        self._sig("fe_neg", None, [u8p, u8p])
        self._sig("pubkey", None, [u8p, u8p])
        self._sig("add_direct", None, [u8p, u8p, u8p])
        self._sig("batch_points", None, [u8p, u8p, C.c_int, u8p])
        self._sig("sha256", None, [u8p, C.c_uint64, u8p])
        self._sig("hash160_comp", None, [C.c_int, u8p, u8p])
        self._sig("hash160_uncomp", None, [u8p, u8p])
        self._sig("eth_addr", None, [u8p, u8p])
        self._sig("xxh64", C.c_uint64, [u8p, C.c_uint64, C.c_uint64])
        self._sig("bloom_new", C.c_void_p, [C.c_uint64])
        self._sig("bloom_free", None, [C.c_void_p])
        self._sig("bloom_desc", None, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)])
        self._sig("bloom_data", C.POINTER(C.c_uint8), [C.c_void_p])
        self._sig("bloom_add", C.c_int, [C.c_void_p, u8p, C.c_int])
        self._sig("bloom_check", C.c_int, [C.c_void_p, u8p, C.c_int])

    def _sig(self, name, res, args):
        f = getattr(self.lib, self.p + name)
        f.restype = res
        f.argtypes = args
        setattr(self, "_" + name, f)

    # ---- primitives -------------------------------------------------------------------------
    def fe_mul(self, a, b):
        o = C.create_string_buffer(32); self._fe_mul(be32(a), be32(b), o); return int.from_bytes(o.raw, "big")

    def fe_add(self, a, b):
        o = C.create_string_buffer(32); self._fe_add(be32(a), be32(b), o); return int.from_bytes(o.raw, "big")

    def fe_sub(self, a, b):
        o = C.create_string_buffer(32); self._fe_sub(be32(a), be32(b), o); return int.from_bytes(o.raw, "big")

    def fe_neg(self, a):
        o = C.create_string_buffer(32); self._fe_neg(be32(a), o); return int.from_bytes(o.raw, "big")

    def fe_sqr(self, a):
        o = C.create_string_buffer(32); self._fe_sqr(be32(a), o); return int.from_bytes(o.raw, "big")

    def fe_inv(self, a):
        o = C.create_string_buffer(32); self._fe_inv(be32(a), o); return int.from_bytes(o.raw, "big")

    def pubkey(self, k):
        o = C.create_string_buffer(64); self._pubkey(be32(k), o)
        return int.from_bytes(o.raw[:32], "big"), int.from_bytes(o.raw[32:], "big")

    def add_direct(self, a, b):
        o = C.create_string_buffer(64)
        self._add_direct(be32(a[0]) + be32(a[1]), be32(b[0]) + be32(b[1]), o)
        return int.from_bytes(o.raw[:32], "big"), int.from_bytes(o.raw[32:], "big")

    def batch_points(self, base, stride, with_y=True):
        o = C.create_string_buffer(1024 * 64)
        self._batch_points(be32(base), be32(stride), 1 if with_y else 0, o)
        return o.raw

    def sha256(self, data: bytes):
        o = C.create_string_buffer(32); self._sha256(data, len(data), o); return o.raw

    def hash160_comp(self, prefix, x):
        o = C.create_string_buffer(20); self._hash160_comp(prefix, be32(x), o); return o.raw

    def hash160_uncomp(self, x, y):
        o = C.create_string_buffer(20); self._hash160_uncomp(be32(x) + be32(y), o); return o.raw

    def eth_addr(self, x, y):
        o = C.create_string_buffer(20); self._eth_addr(be32(x) + be32(y), o); return o.raw

    def xxh64(self, data: bytes, seed: int):
        return self._xxh64(data, len(data), seed)

    # ---- bloom ------------------------------------------------------------------------------
    def bloom_new(self, entries):
        h = self._bloom_new(entries)
        if not h:
            raise ValueError("bloom_init2 rejected entries=%d" % entries)
        return h

    def bloom_free(self, h):
        self._bloom_free(h)

    def bloom_desc(self, h):
        e, b, by, hs = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint32()
        self._bloom_desc(h, C.byref(e), C.byref(b), C.byref(by), C.byref(hs))
        return dict(entries=e.value, bits=b.value, bytes=by.value, hashes=hs.value)

    def bloom_bytes(self, h):
        d = self.bloom_desc(h)
        return C.string_at(self._bloom_data(h), d["bytes"])

    def bloom_add(self, h, data: bytes):
        return self._bloom_add(h, data, len(data))

    def bloom_check(self, h, data: bytes):
        return self._bloom_check(h, data, len(data))


class Oracle(_Lib):
    def __init__(self):
        super().__init__(build_oracle(), "kho_")
        L = self.lib
        u8p = C.c_char_p
        self._sig("ripemd160", None, [u8p, C.c_uint64, u8p])
        self._sig("targets_new", C.c_void_p, [u8p, C.c_uint64])
        self._sig("targets_free", None, [C.c_void_p])
        self._sig("targets_bloom", C.c_void_p, [C.c_void_p])
        self._sig("targets_table", C.POINTER(C.c_uint8), [C.c_void_p, C.POINTER(C.c_uint64)])
        self._sig("searchbinary", C.c_int, [C.c_void_p, u8p])
        self._sig("scan", C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_int, u8p, u8p, C.c_uint64,
                                      C.POINTER(Hit), C.c_uint64, C.c_int])
        self._sig("scan_ex", C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, u8p, u8p, C.c_uint64,
                                         C.POINTER(Hit), C.c_uint64, C.c_int])
        self._sig("bsgs_new", C.c_void_p, [C.c_uint64, C.c_uint32, C.c_int])
        self._sig("bsgs_free", None, [C.c_void_p])
        self._sig("bsgs_params", None, [C.c_void_p] + [C.POINTER(C.c_uint64)] * 5)
        self._sig("bsgs_bloom", C.c_void_p, [C.c_void_p, C.c_int, C.c_int])
        self._sig("bsgs_table", C.POINTER(BpEntry), [C.c_void_p])
        self._sig("bsgs_search", C.c_int, [C.c_void_p, u8p, u8p, u8p, u8p, C.POINTER(C.c_uint64),
                                           C.POINTER(C.c_uint64)])
        self._sig("b58tobin", C.c_int, [u8p, C.POINTER(C.c_uint64), C.c_char_p, C.c_uint64])
        self._sig("addvanity", C.c_int, [C.c_char_p, u8p, u8p, C.c_int, C.POINTER(C.c_int)])
        self._sig("targets_new_vanity", C.c_void_p, [u8p, u8p, C.c_uint64, C.c_int])
        self._sig("bsgs_search_ex", C.c_int, [C.c_void_p, u8p, u8p, u8p, C.c_int, u8p, C.POINTER(C.c_uint64),
                                              C.POINTER(C.c_uint64)])

    def ripemd160(self, data: bytes):
        o = C.create_string_buffer(20); self._ripemd160(data, len(data), o); return o.raw

    # targets / scan
    def targets_new(self, raw20: bytes):
        assert len(raw20) % 20 == 0
        return self._targets_new(raw20, len(raw20) // 20)

    def b58tobin(self, text: bytes, binsz: int):
        """(ok, reported size, buffer) of the reference-style decoder (base58/base58.c:39)"""
        buf = C.create_string_buffer(binsz)
        n = C.c_uint64(binsz)
        ok = self._b58tobin(buf, C.byref(n), text, len(text))
        return bool(ok), n.value, buf.raw

    def addvanity(self, prefixes):
        """addvanity (keyhunt.cpp:6739) over a list of base58 prefixes -> (A limits, B limits, min_bytes, per-prefix r)"""
        A, B, counts = b"", b"", []
        mn = C.c_int(999999)
        for p in prefixes:
            a, b = C.create_string_buffer(20 * 16), C.create_string_buffer(20 * 16)
            r = self._addvanity(p.encode() if isinstance(p, str) else p, a, b, 16, C.byref(mn))
            counts.append(r)
            A += a.raw[:20 * r]; B += b.raw[:20 * r]
        return A, B, mn.value, counts

    def targets_new_vanity(self, A: bytes, B: bytes, min_bytes: int):
        return self._targets_new_vanity(A, B, len(A) // 20, min_bytes)

    def targets_free(self, t):
        self._targets_free(t)

    def targets_table(self, t):
        n = C.c_uint64()
        p = self._targets_table(t, C.byref(n))
        return C.string_at(p, n.value * 20)

    def targets_bloom(self, t):
        return self._targets_bloom(t)

    def searchbinary(self, t, rec: bytes):
        return self._searchbinary(t, rec)

    def scan(self, t, mode, crypto, search, start, stride, n_points, nthreads=8, max_hits=4096, endo=False):
        hits = (Hit * max_hits)()
        n = self._scan_ex(t, mode, crypto, search, 1 if endo else 0, be32(start), be32(stride), n_points, hits, max_hits, nthreads)
        if n < 0:
            raise ValueError("kho_scan: n_points must be a multiple of 1024")
        out = []
        for i in range(min(n, max_hits)):
            h = hits[i]
            out.append(dict(key=int.from_bytes(bytes(h.key_be), "big"), matched=bytes(h.matched),
                            kind=int(h.kind), index=int(h.index), variant=int(h.pad[0])))
        return out

    # bsgs
    def bsgs_new(self, n, k, nthreads=8):
        h = self._bsgs_new(n, k, nthreads)
        if not h:
            raise ValueError("invalid bsgs n/k")
        return h

    def bsgs_free(self, b):
        self._bsgs_free(b)

    def bsgs_params(self, b):
        v = [C.c_uint64() for _ in range(5)]
        self._bsgs_params(b, *[C.byref(x) for x in v])
        return dict(zip(["n", "m", "m2", "m3", "aux"], [x.value for x in v]))

    def bsgs_bloom(self, b, tier, shard):
        return self._bsgs_bloom(b, tier, shard)

    def bsgs_table(self, b):
        m3 = self.bsgs_params(b)["m3"]
        return C.string_at(self._bsgs_table(b), m3 * 16)

    def bsgs_search(self, b, pub, start, end, base_check=False):
        o = C.create_string_buffer(32)
        gs, pos = C.c_uint64(), C.c_uint64()
        r = self._bsgs_search_ex(b, be32(pub[0]) + be32(pub[1]), be32(start), be32(end), int(base_check), o, C.byref(gs), C.byref(pos))
        return (int.from_bytes(o.raw, "big") if r else None), gs.value, pos.value


class RefHarness(_Lib):
    """The reference's own object code (only where oracle/_ref/libkh_ref.so was built)."""

    def __init__(self):
        super().__init__(REF_SO, "khr_")
        self._sig("hash160_scalar", None, [C.c_int, C.c_char_p, C.c_char_p])
        self._sig("b58tobin", C.c_int, [C.c_char_p, C.POINTER(C.c_uint64), C.c_char_p, C.c_uint64])
        self._sig("sizeof_bloom", C.c_int, [])
        self.lib.khr_init()

    def b58tobin(self, text: bytes, binsz: int):
        buf = C.create_string_buffer(binsz)
        n = C.c_uint64(binsz)
        ok = self._b58tobin(buf, C.byref(n), text, len(text))
        return bool(ok), n.value, buf.raw

    def hash160_scalar(self, compressed, x, y):
        o = C.create_string_buffer(20); self._hash160_scalar(1 if compressed else 0, be32(x) + be32(y), o); return o.raw


def have_ref_harness():
    return os.path.exists(REF_SO)


def have_ref_binary():
    return os.path.exists(REF_BIN)
