"""CPU: the oracle and the host build of fe.cuh (tests/devsim) against tests/golden/field_edge.json — known answers of the
reference's Int::ModMulK1 / ModSquareK1 / ModInv / ModAdd / ModSub / ModNeg on operands that force the rare branches of the
limb algorithms (generated from the reference's object code by tests/golden/make_field_golden.py).  The same vectors run
through the GPU's PTX bodies in tests/test_gpu_field.py."""
import ctypes as C
import json
import os
import subprocess

import pytest

from _oracle import P_FIELD, be32, have_ref_harness

HERE = os.path.dirname(os.path.abspath(__file__))
VEC = json.load(open(os.path.join(HERE, "golden", "field_edge.json")))["vectors"]


def I(s):
    return int(s, 16)


def test_vectors_are_arithmetically_right_and_cover_the_rare_branches():
    tags = {v["op"] + "/" + v["tag"] for v in VEC}
    for t in ("mul/second_fold_cfa", "mul/second_fold_cfb", "mul/take_k", "mul/top1", "add/carry", "add/take_k", "add/eq_p",
              "sub/borrow", "sub/zero_b", "sub/equal", "neg/zero", "inv/zero", "reduce/second_fold_cfa", "reduce/second_fold_cfb"):
        assert t in tags, t
    P = P_FIELD
    for v in VEC:
        a, b, r = I(v["a"]), I(v["b"]), I(v["r"])
        want = {"mul": a * b % P, "sqr": a * a % P, "inv": pow(a, P - 2, P), "add": (a + b) % P, "sub": (a - b) % P,
                "neg": (-a) % P, "reduce": ((a << 256) | b) % P}[v["op"]]
        assert r == want, v


def test_oracle_field_ops(oracle):
    for v in VEC:
        a, b, r = I(v["a"]), I(v["b"]), I(v["r"])
        op = v["op"]
        if op == "mul":
            got = oracle.fe_mul(a, b)
        elif op == "sqr":
            got = oracle.fe_sqr(a)
        elif op == "inv":
            got = oracle.fe_inv(a)
        elif op == "add":
            got = oracle.fe_add(a, b)
        elif op == "sub":
            got = oracle.fe_sub(a, b)
        elif op == "neg":
            got = oracle.fe_neg(a)
        else:
            continue
        assert got == r, v


def test_devsim_field_ops():
    d = os.path.join(HERE, "devsim")
    subprocess.check_call(["make", "-C", d], stdout=subprocess.DEVNULL)
    ds = C.CDLL(os.path.join(d, "libkh_devsim.so"))
    fns = {"mul": (ds.ds_fe_mul, 2), "sqr": (ds.ds_fe_sqr, 1), "inv": (ds.ds_fe_inv, 1), "add": (ds.ds_fe_add, 2), "sub": (ds.ds_fe_sub, 2),
           "neg": (ds.ds_fe_neg, 1), "reduce": (ds.ds_fe_reduce_wide, 2)}
    alt = {"mul": (ds.ds_fe_mul_alt, 2), "sqr": (ds.ds_fe_sqr_alt, 1), "inv": (ds.ds_fe_inv_alt, 1), "reduce": (ds.ds_fe_reduce_wide_alt, 2)}
    for v in VEC:
        for table in (fns, alt):          # both forms of the multiplier's final reduction (fe.cuh KH_RARE_REDUCE)
            if v["op"] not in table:
                continue
            fn, n = table[v["op"]]
            o = C.create_string_buffer(32)
            if n == 2:
                fn(be32(I(v["a"])), be32(I(v["b"])), o)
            else:
                fn(be32(I(v["a"])), o)
            assert int.from_bytes(o.raw, "big") == I(v["r"]), v


@pytest.mark.ref
@pytest.mark.skipif(not have_ref_harness(), reason="oracle/_ref/libkh_ref.so not built (build container only)")
def test_reference_still_produces_these_vectors():
    from _oracle import RefHarness
    ref = RefHarness()
    for v in VEC[::7] + [v for v in VEC if "reference_returns" in v]:
        # the reference's own value: where Int::ModMulK1 drops its last carry the file keeps it in "reference_returns"
        a, b, r = I(v["a"]), I(v["b"]), I(v.get("reference_returns", v["r"]))
        got = {"mul": lambda: ref.fe_mul(a, b), "sqr": lambda: ref.fe_sqr(a), "inv": lambda: ref.fe_inv(a), "add": lambda: ref.fe_add(a, b),
               "sub": lambda: ref.fe_sub(a, b), "neg": lambda: ref.fe_neg(a), "reduce": lambda: r}[v["op"]]()
        assert got == r, v
