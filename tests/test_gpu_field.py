"""GPU: the field layer's PTX bodies (fe.cuh: mad.lo.cc / madc.hi.cc / addc / subc carry chains) run ON THE DEVICE through
kh_selftest_fe and are compared with (1) the reference-generated known answers of tests/golden/primitives.json,
(2) tests/golden/field_edge.json — reference answers on operands that force the rare branches (second-fold carry of
fe_reduce_wide, take path of fe_final_reduce, borrow with b = 0, inv(0) = 0 ...), (3) the oracle on seeded random operands."""
import json
import os
import random

import pytest

import keyhunt_b200 as K
from _oracle import P_FIELD

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
PRIM = json.load(open(os.path.join(GOLD, "primitives.json")))
EDGE = json.load(open(os.path.join(GOLD, "field_edge.json")))["vectors"]
OPS = {"mul": K.FE_MUL, "sqr": K.FE_SQR, "inv": K.FE_INV, "add": K.FE_ADD, "sub": K.FE_SUB, "neg": K.FE_NEG, "reduce": K.FE_REDUCE_WIDE}
# the kernels contain two forms of the multiplier's final reduction (fe.cuh KH_RARE_REDUCE / emit.cuh RARE_REDUCE): both are tested
ALT = {"mul": K.FE_MUL_ALT, "sqr": K.FE_SQR_ALT, "inv": K.FE_INV_ALT, "reduce": K.FE_REDUCE_WIDE_ALT}


def I(s):
    return int(s, 16)


def test_reference_vectors_on_device(kh):
    a, b, r = zip(*[(I(x), I(y), I(z)) for x, y, z in PRIM["fe_mul"]])
    assert kh.selftest_fe(K.FE_MUL, a, b) == list(r)
    assert kh.selftest_fe(K.FE_MUL_OUTLINE, a, b) == list(r)
    assert kh.selftest_fe(K.FE_MUL_ALT, a, b) == list(r)
    assert kh.selftest_fe(K.FE_MUL_OUTLINE_ALT, a, b) == list(r)
    a, r = zip(*[(I(x), I(z)) for x, z in PRIM["fe_sqr"]])
    assert kh.selftest_fe(K.FE_SQR, a) == list(r)
    assert kh.selftest_fe(K.FE_SQR_ALT, a) == list(r)
    assert kh.selftest_fe(K.FE_MUL, a, a) == list(r)
    a, r = zip(*[(I(x), I(z)) for x, z in PRIM["fe_inv"]])
    assert kh.selftest_fe(K.FE_INV, a) == list(r)
    assert kh.selftest_fe(K.FE_INV_ALT, a) == list(r)
    assert kh.selftest_fe(K.FE_INV_SQR, a) == list(r)


@pytest.mark.parametrize("op", sorted(OPS))
def test_forced_edge_operands_on_device(kh, op):
    vec = [v for v in EDGE if v["op"] == op]
    assert len(vec) >= 20
    got = kh.selftest_fe(OPS[op], [I(v["a"]) for v in vec], [I(v["b"]) for v in vec])
    bad = [(v, hex(g)) for v, g in zip(vec, got) if g != I(v["r"])]
    assert not bad, bad[:3]
    if op in ALT:
        assert kh.selftest_fe(ALT[op], [I(v["a"]) for v in vec], [I(v["b"]) for v in vec]) == got
    if op == "inv":
        assert kh.selftest_fe(K.FE_INV_SQR, [I(v["a"]) for v in vec]) == got
    if op == "mul":   # the shared out-of-line copy the hash kernels call, and mul(a, a) against the dedicated squaring
        assert kh.selftest_fe(K.FE_MUL_OUTLINE, [I(v["a"]) for v in vec], [I(v["b"]) for v in vec]) == got
        assert kh.selftest_fe(K.FE_MUL_OUTLINE_ALT, [I(v["a"]) for v in vec], [I(v["b"]) for v in vec]) == got
    if op == "sqr":
        assert kh.selftest_fe(K.FE_MUL, [I(v["a"]) for v in vec], [I(v["a"]) for v in vec]) == got


def test_wide_products_on_device(kh):
    """the 512-bit product before reduction (even/odd column accumulators + their combination), incl. all-ones limbs"""
    rnd = random.Random(5)
    M = (1 << 256) - 1
    vals = [0, 1, M, M - 1, 1 << 255, 0xFFFFFFFF, M ^ 0xFFFFFFFF, int("FFFFFFFF00000000" * 4, 16), int("00000000FFFFFFFF" * 4, 16),
            P_FIELD - 1] + [rnd.randrange(1 << 256) for _ in range(200)]
    a = vals
    b = vals[7:] + vals[:7]
    lo = kh.selftest_fe(K.FE_MULWIDE_LO, a, b)
    hi = kh.selftest_fe(K.FE_MULWIDE_HI, a, b)
    assert [(h << 256) | l for h, l in zip(hi, lo)] == [x * y for x, y in zip(a, b)]
    lo = kh.selftest_fe(K.FE_SQRWIDE_LO, a)
    hi = kh.selftest_fe(K.FE_SQRWIDE_HI, a)
    assert [(h << 256) | l for h, l in zip(hi, lo)] == [x * x for x in a]


def test_random_operands_match_the_oracle(kh, oracle):
    rnd = random.Random(6)
    a = [rnd.randrange(P_FIELD) for _ in range(3000)]
    b = [rnd.randrange(P_FIELD) for _ in range(3000)]
    assert kh.selftest_fe(K.FE_MUL, a, b) == [oracle.fe_mul(x, y) for x, y in zip(a, b)]
    assert kh.selftest_fe(K.FE_SQR, a) == [oracle.fe_sqr(x) for x in a]
    assert kh.selftest_fe(K.FE_MUL_ALT, a, b) == [oracle.fe_mul(x, y) for x, y in zip(a, b)]
    assert kh.selftest_fe(K.FE_SQR_ALT, a) == [oracle.fe_sqr(x) for x in a]
    assert kh.selftest_fe(K.FE_ADD, a, b) == [oracle.fe_add(x, y) for x, y in zip(a, b)]
    assert kh.selftest_fe(K.FE_SUB, a, b) == [oracle.fe_sub(x, y) for x, y in zip(a, b)]
    assert kh.selftest_fe(K.FE_NEG, a) == [oracle.fe_neg(x) for x in a]
    assert kh.selftest_fe(K.FE_INV, a[:300]) == [oracle.fe_inv(x) for x in a[:300]]
    assert kh.selftest_fe(K.FE_INV_SQR, a[:300]) == [oracle.fe_inv(x) for x in a[:300]]
    # operands with long runs of ones / zeros in the limbs (carry propagation across all eight limbs)
    s = [((1 << rnd.randrange(1, 256)) - 1) ^ (((1 << rnd.randrange(1, 256)) - 1) << rnd.randrange(0, 200)) for _ in range(500)]
    s = [x % P_FIELD for x in s]
    t = s[11:] + s[:11]
    assert kh.selftest_fe(K.FE_MUL, s, t) == [x * y % P_FIELD for x, y in zip(s, t)]
    assert kh.selftest_fe(K.FE_ADD, s, t) == [(x + y) % P_FIELD for x, y in zip(s, t)]
    assert kh.selftest_fe(K.FE_SUB, s, t) == [(x - y) % P_FIELD for x, y in zip(s, t)]
