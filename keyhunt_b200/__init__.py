"""keyhunt_b200 — B200-native key-range search behind keyhunt's worker semantics.

This module is a thin ctypes binding over the C ABI in ``include/keyhunt_b200.h``
(``keyhunt_b200/libkh_b200.so``, hand-written sm_100a kernels).  It mirrors the reference's worker
surface: a *target set* (what ``readFileAddress`` builds: bloom + sorted table, keyhunt.cpp:7033), a
*scan* of a key sub-range (what one ``thread_process`` chunk does, keyhunt.cpp:3265) and the BSGS
*build* / *search* pair (``thread_bPload`` keyhunt.cpp:5284, ``thread_process_bsgs`` :4549).

There is NO CPU path: importing works anywhere (so build checks can run), but creating a
:class:`KeyHunt` raises unless the CUDA library loads and a GPU is present.
"""
import ctypes as C
import os

__all__ = ["KeyHunt", "KhError", "Hit", "KeyInfo", "BloomDesc", "BsgsDesc", "Stats", "load_library", "LIB_PATH",
           "MODE_XPOINT", "MODE_ADDRESS", "MODE_BSGS", "MODE_RMD160", "MODE_VANITY", "CRYPTO_BTC", "CRYPTO_ETH",
           "SEARCH_UNCOMPRESS", "SEARCH_COMPRESS", "SEARCH_BOTH",
           "HIT_COMP02", "HIT_COMP03", "HIT_UNCOMP", "HIT_ETH", "HIT_XPOINT", "N_ORDER", "parse_targets"]

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KH_B200_LIB") or os.path.join(_HERE, "libkh_b200.so")   # env override: A/B builds when tuning

# keyhunt.cpp:76-90
MODE_XPOINT, MODE_ADDRESS, MODE_BSGS, MODE_RMD160 = 0, 1, 2, 3
MODE_VANITY = 6
CRYPTO_BTC, CRYPTO_ETH = 1, 2
SEARCH_UNCOMPRESS, SEARCH_COMPRESS, SEARCH_BOTH = 0, 1, 2
HIT_COMP02, HIT_COMP03, HIT_UNCOMP, HIT_ETH, HIT_XPOINT = 0, 1, 2, 3, 4
N_ORDER = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141

EXPORTS = ["kh_create", "kh_destroy", "kh_last_error", "kh_set_option", "kh_bloom_params", "kh_set_targets",
           "kh_get_bloom", "kh_get_table", "kh_scan", "kh_poll_hits", "kh_derive", "kh_bsgs_build",
           "kh_bsgs_describe", "kh_bsgs_export", "kh_bsgs_import", "kh_bsgs_digest", "kh_bsgs_search", "kh_get_stats",
           "kh_device_info", "kh_int_peak", "kh_set_vanity", "kh_pipe_peak", "kh_hash_peak", "kh_selftest_fe"]

# kh_selftest_fe ops (include/keyhunt_b200.h)
FE_MUL, FE_SQR, FE_INV, FE_ADD, FE_SUB, FE_NEG, FE_MUL_OUTLINE = 0, 1, 2, 3, 4, 5, 6
FE_MULWIDE_LO, FE_MULWIDE_HI, FE_SQRWIDE_LO, FE_SQRWIDE_HI, FE_REDUCE_WIDE = 7, 8, 9, 10, 11
FE_MUL_ALT, FE_SQR_ALT, FE_INV_ALT, FE_MUL_OUTLINE_ALT, FE_REDUCE_WIDE_ALT, FE_INV_SQR = 12, 13, 14, 15, 16, 17     # the other final-reduction form


class KhError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("kh error %d: %s" % (code, msg))
        self.code = code


class BloomDesc(C.Structure):
    _fields_ = [("entries", C.c_uint64), ("bits", C.c_uint64), ("bytes", C.c_uint64), ("hashes", C.c_uint32),
                ("pad", C.c_uint32)]

    def as_dict(self):
        return dict(entries=self.entries, bits=self.bits, bytes=self.bytes, hashes=self.hashes)


class _Hit(C.Structure):
    _fields_ = [("key_be", C.c_uint8 * 32), ("pub_x", C.c_uint8 * 32), ("pub_y", C.c_uint8 * 32),
                ("matched", C.c_uint8 * 20), ("kind", C.c_uint8), ("pad", C.c_uint8 * 3), ("index", C.c_uint64)]


class _KeyInfo(C.Structure):
    _fields_ = [("pub_x", C.c_uint8 * 32), ("pub_y", C.c_uint8 * 32), ("h160_comp", C.c_uint8 * 20),
                ("h160_uncomp", C.c_uint8 * 20), ("eth", C.c_uint8 * 20), ("pad", C.c_uint8 * 4)]


class BsgsDesc(C.Structure):
    _fields_ = [("n", C.c_uint64), ("m", C.c_uint64), ("m2", C.c_uint64), ("m3", C.c_uint64), ("aux", C.c_uint64),
                ("tier", BloomDesc * 3)]


class Stats(C.Structure):
    _fields_ = [("walk_ms", C.c_double), ("setup_ms", C.c_double), ("aux_ms", C.c_double),
                ("walk_launches", C.c_uint64), ("other_launches", C.c_uint64), ("points", C.c_uint64),
                ("walker_threads", C.c_uint64), ("tier1_positives", C.c_uint64), ("collapsed_batches", C.c_uint64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class Hit:
    """One reported key (the reference's writekey record, keyhunt.cpp:6891)."""
    __slots__ = ("key", "pub_x", "pub_y", "matched", "kind", "index", "variant")

    def __init__(self, h):
        self.key = int.from_bytes(bytes(h.key_be), "big")
        self.pub_x = int.from_bytes(bytes(h.pub_x), "big")
        self.pub_y = int.from_bytes(bytes(h.pub_y), "big")
        self.matched = bytes(h.matched)
        self.kind = int(h.kind)
        self.index = int(h.index)
        self.variant = int(h.pad[0])      # -e: the reference's candidate index l

    def __repr__(self):
        return "Hit(key=%x kind=%d index=%d matched=%s)" % (self.key, self.kind, self.index, self.matched.hex())


class KeyInfo:
    __slots__ = ("pub_x", "pub_y", "h160_comp", "h160_uncomp", "eth")

    def __init__(self, k):
        self.pub_x = int.from_bytes(bytes(k.pub_x), "big")
        self.pub_y = int.from_bytes(bytes(k.pub_y), "big")
        self.h160_comp = bytes(k.h160_comp)
        self.h160_uncomp = bytes(k.h160_uncomp)
        self.eth = bytes(k.eth)


_lib = None


def load_library(path=None):
    """dlopen libkh_b200.so and declare every prototype of include/keyhunt_b200.h.  Needs no GPU."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if path is None and os.environ.get("KH_B200_LIB"):
        path = p                                          # A/B build named by the environment: tolerate missing newest symbols
    if not os.path.exists(p):
        raise OSError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(nvcc, sm_100a). keyhunt_b200 has no CPU fallback." % p)
    L = C.CDLL(p)
    vp, u8p, u64 = C.c_void_p, C.c_char_p, C.c_uint64
    L.kh_create.argtypes = [C.POINTER(vp), C.c_int]
    L.kh_destroy.argtypes = [vp]; L.kh_destroy.restype = None
    L.kh_last_error.argtypes = [vp]; L.kh_last_error.restype = C.c_char_p
    L.kh_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.kh_bloom_params.argtypes = [u64, C.POINTER(BloomDesc)]
    L.kh_set_targets.argtypes = [vp, C.c_int, C.c_int, C.c_int, u8p, u64, C.POINTER(BloomDesc), u8p]
    L.kh_get_bloom.argtypes = [vp, C.POINTER(BloomDesc), vp, u64]
    L.kh_get_table.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.kh_scan.argtypes = [vp, u8p, u8p, u64]
    L.kh_poll_hits.argtypes = [vp, C.POINTER(_Hit), C.c_int, C.POINTER(C.c_int)]
    L.kh_derive.argtypes = [vp, u8p, u64, C.POINTER(_KeyInfo)]
    L.kh_bsgs_build.argtypes = [vp, u64, C.c_uint32]
    L.kh_bsgs_describe.argtypes = [vp, C.POINTER(BsgsDesc)]
    L.kh_bsgs_export.argtypes = [vp, C.c_int, C.c_int, vp, u64]
    L.kh_bsgs_import.argtypes = [vp, C.c_int, C.c_int, u8p, u64]
    L.kh_bsgs_search.argtypes = [vp, u8p, u8p, u8p, u8p, C.POINTER(C.c_int)]
    L.kh_get_stats.argtypes = [vp, C.POINTER(Stats), C.c_int]
    L.kh_device_info.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(u64)]
    L.kh_int_peak.argtypes = [vp, C.POINTER(C.c_double)]
    if path is None or hasattr(L, "kh_selftest_fe"):      # (an explicitly named older A/B build may lack the newest entry points)
        L.kh_pipe_peak.argtypes = [vp, C.POINTER(C.c_double)]
        L.kh_hash_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
        L.kh_selftest_fe.argtypes = [vp, C.c_int, u8p, u8p, u64, vp]
    if path is None or hasattr(L, "kh_bsgs_digest"):
        L.kh_bsgs_digest.argtypes = [vp, C.c_int, C.POINTER(u64)]
    for name in EXPORTS:
        if name not in ("kh_destroy", "kh_last_error") and (path is None or hasattr(L, name)):
            getattr(L, name).restype = C.c_int
    if path is None or path == LIB_PATH:
        _lib = L
    return L


def bloom_params(entries):
    d = BloomDesc()
    rc = load_library().kh_bloom_params(entries, C.byref(d))
    if rc:
        raise KhError(rc, "bloom_init2 rejects entries=%d" % entries)
    return d


def _be32(v):
    return int(v).to_bytes(32, "big")


class KeyHunt:
    """One GPU context (one per device, like one process of the reference per box)."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.kh_create(C.byref(h), device)
        if rc:
            raise KhError(rc, "kh_create(device=%d) failed: no usable CUDA device (there is no CPU fallback)" % device)
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.kh_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise KhError(rc, self._lib.kh_last_error(self._h).decode())

    def set_option(self, name, value):
        self._ck(self._lib.kh_set_option(self._h, name.encode(), int(value)))

    def device_info(self):
        name = C.create_string_buffer(256)
        sm = C.c_int()
        mem = C.c_uint64()
        self._ck(self._lib.kh_device_info(self._h, name, 256, C.byref(sm), C.byref(mem)))
        return dict(name=name.value.decode(), sm_count=sm.value, hbm_bytes=mem.value)

    # ---- scan modes ---------------------------------------------------------------------------
    def set_targets(self, mode, records20, crypto=CRYPTO_BTC, search=SEARCH_COMPRESS, desc=None, bloom_bits=None):
        """records20: bytes, N x 20 (hash160 / ETH address / first 20 bytes of X), any order."""
        if len(records20) % 20:
            raise ValueError("records20 must be a multiple of 20 bytes")
        self._ck(self._lib.kh_set_targets(self._h, mode, crypto, search, records20, len(records20) // 20,
                                          C.byref(desc) if desc is not None else None, bloom_bits))

    def set_vanity(self, limits_a, limits_b, search=SEARCH_COMPRESS):
        """-m vanity: limits_a / limits_b = N x 20 bytes, the interval pairs addvanity (keyhunt.cpp:6739) produced."""
        if len(limits_a) != len(limits_b) or len(limits_a) % 20:
            raise ValueError("limits must be two equal multiples of 20 bytes")
        self._ck(self._lib.kh_set_vanity(self._h, search, limits_a, limits_b, len(limits_a) // 20))

    def get_bloom(self):
        d = BloomDesc()
        self._ck(self._lib.kh_get_bloom(self._h, C.byref(d), None, 0))
        buf = C.create_string_buffer(d.bytes)
        self._ck(self._lib.kh_get_bloom(self._h, C.byref(d), buf, d.bytes))
        return d, buf.raw

    def get_table(self):
        n = C.c_uint64()
        self._ck(self._lib.kh_get_table(self._h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(max(1, n.value * 20))
        self._ck(self._lib.kh_get_table(self._h, buf, n.value, C.byref(n)))
        return buf.raw[:n.value * 20]

    def scan(self, start, n_points, stride=1):
        """Scan keys start + i*stride for i < n_points (multiple of 1024). Blocking."""
        self._ck(self._lib.kh_scan(self._h, _be32(start), _be32(stride), n_points))

    def poll_hits(self, max_hits=4096):
        """Drains the hits of the scans since the last poll.  If the device hit buffer overflowed (KH_EOVERFLOW) the hits
        that were kept are still returned: the overflow is raised as KhError carrying them in ``.hits``."""
        out = []
        overflow = None
        while True:
            arr = (_Hit * max_hits)()
            n = C.c_int()
            rc = self._lib.kh_poll_hits(self._h, arr, max_hits, C.byref(n))
            out.extend(Hit(arr[i]) for i in range(n.value))
            if rc == -5:                       # KH_EOVERFLOW: kh_poll_hits has already handed over (and erased) these hits
                overflow = KhError(rc, self._lib.kh_last_error(self._h).decode())
            elif rc:
                self._ck(rc)
            if n.value < max_hits:
                break
        if overflow is not None:
            overflow.hits = out
            raise overflow
        return out

    def derive(self, keys):
        keys = list(keys)
        arr = (_KeyInfo * max(1, len(keys)))()
        self._ck(self._lib.kh_derive(self._h, b"".join(_be32(k) for k in keys), len(keys), arr))
        return [KeyInfo(arr[i]) for i in range(len(keys))]

    # ---- bsgs ---------------------------------------------------------------------------------
    def bsgs_build(self, n, k=1):
        self._ck(self._lib.kh_bsgs_build(self._h, n, k))

    def bsgs_describe(self):
        d = BsgsDesc()
        self._ck(self._lib.kh_bsgs_describe(self._h, C.byref(d)))
        return d

    def bsgs_export(self, tier, shard=0):
        d = self.bsgs_describe()
        size = d.m3 * 16 if tier == 0 else d.tier[tier - 1].bytes
        buf = C.create_string_buffer(max(1, size))
        self._ck(self._lib.kh_bsgs_export(self._h, tier, shard, buf, size))
        return buf.raw[:size]

    def bsgs_digest(self, tier):
        """device-side digest of the bP table (0), a whole bloom tier (1..3) or the baby-point prefix bitmap (4)"""
        out = C.c_uint64()
        self._ck(self._lib.kh_bsgs_digest(self._h, tier, C.byref(out)))
        return out.value

    def bsgs_import(self, tier, shard, data):
        self._ck(self._lib.kh_bsgs_import(self._h, tier, shard, data, len(data)))

    def bsgs_search(self, pub, start, end):
        """pub = (x, y) ints. Returns the private key or None."""
        out = C.create_string_buffer(32)
        found = C.c_int()
        self._ck(self._lib.kh_bsgs_search(self._h, _be32(pub[0]) + _be32(pub[1]), _be32(start), _be32(end), out,
                                          C.byref(found)))
        return int.from_bytes(out.raw, "big") if found.value else None

    def int_peak(self):
        """measured integer-pipe peaks, thread-ops/s over the whole chip"""
        arr = (C.c_double * 6)()
        self._ck(self._lib.kh_int_peak(self._h, arr))
        return dict(zip(["iadd3", "lop3", "shf", "imad", "imad_wide", "lop3_imad_mix"], [float(x) for x in arr]))

    def pipe_peak(self):
        """more measured pipe rates (thread-ops/s, whole chip): the evidence behind the choice of multiplier"""
        arr = (C.c_double * 16)()
        self._ck(self._lib.kh_pipe_peak(self._h, arr))
        return dict(zip(["imad_wide_nocarry", "imad_hi", "dfma", "dadd", "dfma_plus_imad_wide", "imad_wide_plus_iadd3", "ffma", "imad_wide_in_walk_mix", "imad_wide_nocarry_plus_lop3",
                         "imad_wide_carryout_plus_iadd3x", "imad_wide_nocarry_plus_iadd3", "imad_wide_2link_chain_plus_iadd3x"],
                        [float(x) for x in arr]))

    def hash_peak(self, blocks_per_sm=2):
        arr = (C.c_double * 2)()
        self._ck(self._lib.kh_hash_peak(self._h, blocks_per_sm, arr))
        return {"sha256_compressions_per_s": float(arr[0]), "ripemd160_blocks_per_s": float(arr[1])}

    def selftest_fe(self, op, a, b=None):
        """one field operation per element ON THE DEVICE (fe.cuh's PTX bodies); a, b: lists of ints -> list of ints"""
        a = list(a)
        b = list(b) if b is not None else [0] * len(a)
        if len(a) != len(b):
            raise ValueError("operand lists differ in length")
        out = C.create_string_buffer(max(1, 32 * len(a)))
        self._ck(self._lib.kh_selftest_fe(self._h, op, b"".join(_be32(x) for x in a), b"".join(_be32(x) for x in b), len(a), out))
        return [int.from_bytes(out.raw[32 * i:32 * i + 32], "big") for i in range(len(a))]

    def stats(self, reset=False):
        s = Stats()
        self._ck(self._lib.kh_get_stats(self._h, C.byref(s), 1 if reset else 0))
        return s.as_dict()


# ---- host-side helpers mirroring the reference's target-file parsers ------------------------------
_B58 = "123456789ABCDEFGHJKLMNPQRSTUVWXYZabcdefghijkmnopqrstuvwxyz"


def _b58decode25(s):
    v = 0
    for ch in s:
        v = v * 58 + _B58.index(ch)
    return v.to_bytes(25, "big")


def parse_targets(lines, mode, crypto=CRYPTO_BTC):
    """20-byte records from a keyhunt ``-f`` file, following forceReadFileAddress (keyhunt.cpp:7239),
    forceReadFileAddressEth (:7312) and forceReadFileXPoint (:7392).  Invalid lines are skipped."""
    out = []
    for raw in lines:
        s = raw.strip(" \t\r\n")
        if not s:
            continue
        tok = s.split()[0]
        try:
            if mode == MODE_XPOINT:
                if len(tok) == 64:
                    out.append(bytes.fromhex(tok)[:20])
                elif len(tok) == 66:
                    out.append(bytes.fromhex(tok[2:])[:20])
                elif len(tok) == 130:
                    # reference quirk (keyhunt.cpp:7463-7466): table gets rawvalue+2, bloom rawvalue+0,
                    # so such targets can never hit; we keep the table bytes.
                    out.append(bytes.fromhex(tok)[2:22])
            elif crypto == CRYPTO_ETH:
                if len(s) == 40:
                    out.append(bytes.fromhex(s))
                elif len(s) == 42:
                    out.append(bytes.fromhex(s[2:]))
            else:
                if len(s) == 40 and all(c in "0123456789abcdefABCDEF" for c in s):
                    out.append(bytes.fromhex(s))
                elif 20 < len(s) < 40 and all(c in _B58 for c in s):
                    out.append(_b58decode25(s)[1:21])
        except (ValueError, OverflowError):
            continue
    return b"".join(out)
