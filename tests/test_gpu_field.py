"""GPU parity: device Fq/Fr/G1/G2 arithmetic vs the Python-int oracle, bit-exact (integer work)."""
import ctypes as C
import random

import numpy as np
import pytest

from oracle import bn254 as bn
from simple_zk_rollups_b200 import _lib
from helpers import pack, unpack, pack_g1, pack_g2, unpack_g1, unpack_g2

pytestmark = pytest.mark.gpu
MONT = 1 << 256


def edge_values(p, rng, n):
    ed = [0, 1, 2, p - 1, p - 2, (p + 1) // 2, MONT % p, (MONT * MONT) % p, (1 << 253), (1 << 224) - 1]
    return ed + [rng.randrange(p) for _ in range(n - len(ed))]


@pytest.mark.parametrize("field,p", [(0, bn.Q), (1, bn.R)])
def test_field_ops(zctx, field, p):
    L = _lib.lib()
    rng = random.Random(1234 + field)
    n = 100000
    a = edge_values(p, rng, n)
    b = list(reversed(edge_values(p, rng, n)))
    rng.shuffle(b)
    # operands are interpreted as Montgomery residues: x*y*R^-1
    rinv = pow(MONT, -1, p)
    A, B = pack(a), pack(b)
    out = np.zeros_like(A)
    exp = {0: [x * y * rinv % p for x, y in zip(a, b)],
           1: [(x + y) % p for x, y in zip(a, b)],
           2: [(x - y) % p for x, y in zip(a, b)],
           3: [x * x * rinv % p for x in a],
           5: [x * MONT % p for x in a],
           6: [x * rinv % p for x in a],
           # sums of products with one reduction (fp.cuh mont_mul2_raw / mont_mul4_raw; the lazily reduced additions use them)
           7: [(x * y + (x + y) * (x - y)) * rinv % p for x, y in zip(a, b)],
           8: [(x * y - (x + y) * (x - y)) * rinv % p for x, y in zip(a, b)],
           9: [(x * y + (x + y) * (x - y) + x * (x - y) + (x + y) * y) * rinv % p for x, y in zip(a, b)],
           10: [x * x * rinv % p for x in a]}                      # the dedicated squaring (mont_sqr_raw)
    for op, e in exp.items():
        _lib.check(L.zkr_test_field_op(zctx, field, op, _lib.buf_ptr(A), _lib.buf_ptr(B), _lib.buf_ptr(out), n))
        got = unpack(out)
        bad = [i for i in range(n) if got[i] != e[i]]
        assert not bad, "field %d op %d first mismatch at %d" % (field, op, bad[0])
    # inverse on a smaller batch (Fermat, ~380 modmuls each); Montgomery: inv(xR) = x^-1 R
    m = 2000
    nz = [v if v else 1 for v in a[:m]]
    A2 = pack(nz)
    out2 = np.zeros_like(A2)
    _lib.check(L.zkr_test_field_op(zctx, field, 4, _lib.buf_ptr(A2), None, _lib.buf_ptr(out2), m))
    got = unpack(out2)
    for x, g in zip(nz, got):
        assert g == pow(x * rinv % p, -1, p) * MONT % p


def _g1_pts(rng, n):
    fb = bn.fixed_base(1)
    return fb.mul_many([rng.randrange(1, bn.R) for _ in range(n)])


def _g2_pts(rng, n):
    fb = bn.fixed_base(2)
    return fb.mul_many([rng.randrange(1, bn.R) for _ in range(n)])


@pytest.mark.parametrize("group", [1, 2])
def test_curve_ops(zctx, group):
    L = _lib.lib()
    rng = random.Random(99 + group)
    n = 256
    cur = bn.G1 if group == 1 else bn.G2
    gen_pts = _g1_pts if group == 1 else _g2_pts
    pk, up = (pack_g1, unpack_g1) if group == 1 else (pack_g2, unpack_g2)
    P = gen_pts(rng, n)
    Q = gen_pts(rng, n)
    # exceptional cases: Q == P, Q == -P, P infinity, Q infinity
    Q[0] = P[0]
    Q[1] = cur.neg(P[1])
    P[2] = None
    Q[3] = None
    P[4] = None
    Q[4] = None
    Pa, Qa = pk(P), pk(Q)
    out = np.zeros_like(Pa)
    _lib.check(L.zkr_test_curve_op(zctx, group, 0, _lib.buf_ptr(Pa), _lib.buf_ptr(Qa), _lib.buf_ptr(out), n))
    assert up(out) == [cur.add(p, q) for p, q in zip(P, Q)]
    _lib.check(L.zkr_test_curve_op(zctx, group, 1, _lib.buf_ptr(Pa), None, _lib.buf_ptr(out), n))
    assert up(out) == [cur.add(p, p) for p in P]
    _lib.check(L.zkr_test_curve_op(zctx, group, 3, _lib.buf_ptr(Pa), _lib.buf_ptr(Qa), _lib.buf_ptr(out), n))
    assert up(out) == [cur.add(cur.add(p, p), cur.add(q, q)) for p, q in zip(P, Q)]
    # 2P + Q through the mixed addition, plain and lazily reduced (Q == 2P and Q == -2P are its exceptional branches)
    Q2 = list(Q)
    Q2[5] = cur.add(P[5], P[5])
    Q2[6] = cur.neg(cur.add(P[6], P[6]))
    Q2a = pk(Q2)
    want = [cur.add(cur.add(p, p), q) for p, q in zip(P, Q2)]
    for op in (4, 5, 6):
        _lib.check(L.zkr_test_curve_op(zctx, group, op, _lib.buf_ptr(Pa), _lib.buf_ptr(Q2a), _lib.buf_ptr(out), n))
        assert up(out) == want, "curve op %d" % op
    ks = [0, 1, 2, bn.R - 1, bn.R, (1 << 256) - 1] + [rng.randrange(bn.R) for _ in range(n - 6)]
    K = pack(ks)
    _lib.check(L.zkr_test_curve_op(zctx, group, 2, _lib.buf_ptr(Pa), _lib.buf_ptr(K), _lib.buf_ptr(out), n))
    assert up(out) == [cur.mul(p, k) for p, k in zip(P, ks)]


def test_microbench_runs(zctx):
    L = _lib.lib()
    ops = C.c_double()
    ms = C.c_float()
    for which, it in ((0, 20000), (1, 20000), (2, 2000), (3, 500), (12, 500), (13, 200), (14, 200), (15, 500)):
        _lib.check(L.zkr_microbench(zctx, which, it, C.byref(ops), C.byref(ms)))
        assert ops.value > 0
