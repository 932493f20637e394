"""GPU: the BASELINE.json configurations at their own sizes and starts (C2 and C5), the degenerate batches the advisor
pointed at (centre on the hop point; a batch without shared inverse), the range-wrap refusal, and device-memory hygiene."""
import random

import pytest

import keyhunt_b200 as K
from _oracle import N_ORDER, be32

pytestmark = pytest.mark.gpu


def _plant(kh, oracle, start, idxs, field):
    """records of the planted keys: derived on the device AND checked against the oracle's derivation"""
    infos = kh.derive([start + i for i in idxs])
    recs = []
    for j, (i, info) in enumerate(zip(idxs, infos)):
        x, y = oracle.pubkey(start + i)
        assert (info.pub_x, info.pub_y) == (x, y)
        want = {"eth": oracle.eth_addr(x, y), "comp": oracle.hash160_comp(2 + (y & 1), x), "uncomp": oracle.hash160_uncomp(x, y)}
        f = field if field != "both" else ("uncomp" if j % 2 else "comp")
        r = {"eth": info.eth, "comp": info.h160_comp, "uncomp": info.h160_uncomp}[f]
        assert r == want[f]
        recs.append(r)
    return recs


def _idxs(seed, n, count):
    rnd = random.Random(seed)
    s = {0, n - 1}
    while len(s) < count:
        s.add(rnd.randrange(n))
    return sorted(s)


@pytest.mark.parametrize("crypto,field", [(K.CRYPTO_ETH, "eth"), (K.CRYPTO_BTC, "comp")])
def test_c5_config_eth_and_btc_hit_sets(kh, oracle, crypto, field):
    """BASELINE configs[4] as stated: -m address, start 0x10000000000, 1,024 targets of which 16 planted; the ETH and the
    BTC-compress runs are separate (the reference takes one -c per run, keyhunt.cpp:884-888) and so are their hit sets.
    One 2^32-key shard of the 2^40 range (what one step of one GPU scans)."""
    start, n = 0x10000000000, 1 << 32
    idxs = _idxs(50 + crypto, n, 16)
    rnd = random.Random(51)
    planted = _plant(kh, oracle, start, idxs, field)
    recs = planted + [rnd.randbytes(20) for _ in range(1024 - 16)]
    rnd.shuffle(recs)
    kh.set_targets(K.MODE_ADDRESS, b"".join(recs), crypto=crypto, search=K.SEARCH_COMPRESS)
    kh.scan(start, n)
    hits = kh.poll_hits()
    assert sorted((h.index, h.key, h.matched) for h in hits) == sorted((i, start + i, r) for i, r in zip(idxs, planted))
    for h in hits:
        assert h.kind == (K.HIT_ETH if field == "eth" else (K.HIT_COMP02 + (h.pub_y & 1)))
        assert (h.pub_x, h.pub_y) == oracle.pubkey(h.key)
    # the other currency's targets must not match this run (hit sets separate)
    other = _plant(kh, oracle, start, idxs[:4], "comp" if field == "eth" else "eth")
    kh.set_targets(K.MODE_ADDRESS, b"".join(other + [bytes([7]) * 20] * 12), crypto=crypto, search=K.SEARCH_COMPRESS)
    kh.scan(start, 1 << 20)
    assert kh.poll_hits() == []


def test_c2_config_full_size_step(kh, oracle):
    """BASELINE configs[1]: rmd160 -l both, 1,024 hash160 targets (24 planted, both encodings), one full 2^32-key step of
    the 2^36 sweep at its real start; every planted key and nothing else"""
    start, n = 0x2000000000000000, 1 << 32
    idxs = _idxs(2, n, 24)
    rnd = random.Random(3)
    planted = _plant(kh, oracle, start, idxs, "both")
    recs = planted + [rnd.randbytes(20) for _ in range(1000)]
    rnd.shuffle(recs)
    kh.set_targets(K.MODE_RMD160, b"".join(recs), crypto=K.CRYPTO_BTC, search=K.SEARCH_BOTH)
    kh.stats(reset=True)
    kh.scan(start, n)
    st = kh.stats()
    assert sorted((h.index, h.key, h.matched) for h in kh.poll_hits()) == sorted((i, start + i, r) for i, r in zip(idxs, planted))
    assert st["points"] == n and st["collapsed_batches"] == 0 and st["walker_threads"] * 1024 <= n


def test_centre_on_the_hop_point(kh, oracle):
    """start = 512, stride 1: walker T-1 starts with its centre ON W = T*1024*G (the hop's difference is zero).
    Keys of that walker's first, second and third batch must be found, as the reference finds them."""
    kh.set_option("threads_per_sm", 32)
    try:
        info = kh.device_info()
        T = (info["sm_count"] * 32) // 256 * 256
        n = 3 * T * 1024
        idxs = sorted({(T - 1) * 1024 + o for o in (0, 1, 511, 512, 513, 1023)} | {(2 * T - 1) * 1024 + o for o in (0, 512, 1023)} |
                      {(3 * T - 1) * 1024 + 77, 5, n - 1, T * 1024})
        recs = []
        for i in idxs:
            x, _ = oracle.pubkey(512 + i)
            recs.append(be32(x)[:20])
        kh.set_targets(K.MODE_XPOINT, b"".join(recs) + bytes(20))
        kh.stats(reset=True)
        kh.scan(512, n)
        st = kh.stats()
        assert st["walker_threads"] == T and st["collapsed_batches"] == 0
        assert sorted(h.index for h in kh.poll_hits()) == idxs
    finally:
        kh.set_option("threads_per_sm", 4096)


def test_batch_without_shared_inverse(kh, oracle):
    """a range that runs over key 0 (mod n): two batches have a centre +-512*G, no shared inverse exists; like the reference
    (IntGroup::ModInv gives zeros) only the CENTRES of those batches are tested, they are counted, and
    every later batch of the same walkers is right"""
    kh.set_option("threads_per_sm", 32)
    try:
        start, n = N_ORDER - 1024, 1 << 23
        keys = {0: "centre-only batch: not the centre", 512: "centre of batch 0", 1536: "centre of batch 1", 1024 + 5: "collapsed",
                2048: "first good batch", 2048 + 1023: "", (1 << 22) + 17: "", n - 1: ""}
        recs = []
        for i in keys:
            x, _ = oracle.pubkey((start + i) % N_ORDER)
            recs.append(be32(x)[:20])
        kh.set_targets(K.MODE_XPOINT, b"".join(recs) + bytes(20))
        t = oracle.targets_new(b"".join(recs) + bytes(20))
        kh.stats(reset=True)
        kh.scan(start, n)
        st = kh.stats()
        got = sorted((h.index, h.key) for h in kh.poll_hits())
        assert st["collapsed_batches"] == 2
        assert [g[0] for g in got] == [512, 1536, 2048, 2048 + 1023, (1 << 22) + 17, n - 1]
        assert all(k == (start + i) % N_ORDER for i, k in got)
        # the oracle restates the reference on the first 4 batches: same indices
        from _oracle import CRYPTO_BTC, MODE_XPOINT, SEARCH_COMPRESS
        want = sorted(h["index"] for h in oracle.scan(t, MODE_XPOINT, CRYPTO_BTC, SEARCH_COMPRESS, start, 1, 4096))
        oracle.targets_free(t)
        assert want == [g[0] for g in got if g[0] < 4096]
    finally:
        kh.set_option("threads_per_sm", 4096)


def test_range_that_wraps_2_256_is_refused(kh):
    kh.set_targets(K.MODE_XPOINT, bytes(20))
    with pytest.raises(K.KhError) as ei:
        kh.scan(1, 1 << 20, stride=1 << 240)            # start + stride*n_points >= 2^256
    assert ei.value.code == -2
    with pytest.raises(K.KhError):
        kh.scan((1 << 256) - 5, 1024)
    kh.scan(1, 1024, stride=1 << 200)                   # fits
    assert kh.poll_hits() == []


def test_create_build_destroy_returns_device_memory():
    """kh_destroy must release every buffer, the BSGS prefix bitmap included"""
    import torch
    torch.cuda.init()
    free0 = torch.cuda.mem_get_info(0)[0]
    for _ in range(2):
        with K.KeyHunt(0) as k2:
            k2.bsgs_build(1 << 24, 4)
            k2.set_targets(K.MODE_XPOINT, bytes(40))
            k2.scan(1, 1 << 20)
            used = free0 - torch.cuda.mem_get_info(0)[0]
            assert used > (1 << 20)
        free1 = torch.cuda.mem_get_info(0)[0]
        assert abs(free0 - free1) <= (64 << 20), (free0, free1)    # the context's own pools may keep a few MB
