"""The device field / curve arithmetic (csrc/fp.cuh, csrc/ec.cuh) built for the HOST and checked against Python integers.

Under g++ the four carry-chain row primitives of fp.cuh are portable C; everything composed from them is the code the
kernels run: the even / odd CIOS, the multi-product CIOS with one reduction (mont_mul2_raw / mont_mul4_raw), Fq2 in its
Karatsuba and schoolbook-lazy forms, and the XYZZ mixed addition in both forms.  The bound the lazy forms rest on -- the odd
accumulator never carries out of 256 bits for up to four products of operands <= p -- is asserted on every call
(fp_host_carry_lost), including on all-maximal operands."""
import ctypes as C
import os
import random
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import bn254 as bn
from helpers import pack, unpack, pack_g1, pack_g2

HERE = os.path.dirname(os.path.abspath(__file__))
MONT = 1 << 256


def _build(tag, *defines):
    d = tempfile.mkdtemp(prefix="zkr_fphost_")
    so = os.path.join(d, "fp_host_%s.so" % tag)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", *defines, "-o", so,
                           os.path.join(HERE, "csrc", "fp_host.cpp")])
    return C.CDLL(so)


@pytest.fixture(scope="module")
def lib():
    return _build("a", "-DZKR_LAZY_TAIL=0")


@pytest.fixture(scope="module")
def lib_tail():
    return _build("b", "-DZKR_LAZY_TAIL=1")


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def operands(p, rng, n):
    """n operand 4-tuples: the corners first (0, 1, p - 1 and the non-canonical p the lazy negation can produce), then random"""
    corner = [0, 1, 2, p - 1, p - 2, p, (p + 1) // 2, MONT % p, (1 << 253)]
    tup = [(a, b, c, d) for a in (0, 1, p - 1, p) for b in (0, p - 1, p) for c in (0, p - 1, p) for d in (1, p - 1, p)]
    # limb patterns: carries between limbs, the bit the squaring's doubled operand moves from limb i into limb i + 1
    pat = (0, 1, 0x7fffffff, 0x80000000, 0xffffffff, 0xfffffffe, 0x80000001)

    def patterned():
        limbs = [rng.choice(pat) for _ in range(7)] + [rng.choice((0, 1, 0x10000000, 0x30644e71))]
        return sum(l << (32 * i) for i, l in enumerate(limbs))      # top limb below the moduli's: the value is < p
    for _ in range(3000):
        tup.append(tuple(patterned() for _ in range(4)))
    while len(tup) < n:
        tup.append(tuple(rng.choice(corner) if rng.random() < 0.15 else rng.randrange(p) for _ in range(4)))
    return tup[:n]


@pytest.mark.parametrize("field,p", [(0, bn.Q), (1, bn.R)])
def test_field_ops_match_python(lib, field, p):
    rng = random.Random(77 + field)
    n = 20000
    tup = operands(p, rng, n)
    A, B, Cc, D = (pack([t[i] for t in tup]) for i in range(4))
    out = np.zeros_like(A)
    rinv = pow(MONT, -1, p)
    exp = {
        0: lambda a, b, c, d: a * b * rinv % p,
        3: lambda a, b, c, d: (a * b + c * d) * rinv % p,
        4: lambda a, b, c, d: (a * b - c * d) * rinv % p,
        6: lambda a, b, c, d: a * a * rinv % p,
        8: lambda a, b, c, d: a * a * rinv % p,          # the dedicated squaring (operands < 2^254 by contract: see below)
        7: lambda a, b, c, d: (a * b + c * d + a * d + c * b) * rinv % p,
    }
    for op, f in exp.items():
        if op == 8:      # mont_sqr_raw doubles its operand inside 256 bits: canonical operands (< p < 2^254) only
            A8 = pack([t[0] % p for t in tup])
            lib.fp_host_op(field, op, ptr(A8), ptr(B), ptr(Cc), ptr(D), ptr(out), n)
            assert lib.fp_host_carry_lost() == 0
            got = unpack(out)
            bad = [i for i in range(n) if got[i] != f(tup[i][0] % p, 0, 0, 0)]
            assert not bad, "field %d sqr_fast: first mismatch at %d %s" % (field, bad[0], hex(tup[bad[0]][0]))
            continue
        lib.fp_host_op(field, op, ptr(A), ptr(B), ptr(Cc), ptr(D), ptr(out), n)
        assert lib.fp_host_carry_lost() == 0, "op %d: a carry the code calls impossible occurred" % op
        got = unpack(out)
        bad = [i for i in range(n) if got[i] != f(*tup[i])]
        assert not bad, "field %d op %d: first mismatch at %d %s" % (field, op, bad[0], [hex(v) for v in tup[bad[0]]])
    # add / sub / lazy negation are defined on canonical operands
    can = [tuple(v % p for v in t) for t in tup]
    A, B = pack([t[0] for t in can]), pack([t[1] for t in can])
    for op, f in {1: lambda a, b: (a + b) % p, 2: lambda a, b: (a - b) % p, 5: lambda a, b: p - a}.items():
        lib.fp_host_op(field, op, ptr(A), ptr(B), ptr(A), ptr(B), ptr(out), n)
        got = unpack(out)
        bad = [i for i in range(n) if got[i] != f(can[i][0], can[i][1])]
        assert not bad, "field %d op %d: first mismatch at %d" % (field, op, bad[0])


def _f2(rng, p, corner=0.1):
    def one():
        return rng.choice([0, 1, p - 1]) if rng.random() < corner else rng.randrange(p)
    return (one(), one())


def _pack_f2(vals):
    return pack([c for v in vals for c in v])


def _unpack_f2(arr):
    v = unpack(arr)
    return [(v[i], v[i + 1]) for i in range(0, len(v), 2)]


def test_fq2_lazy_forms_match_python(lib):
    p = bn.Q
    rng = random.Random(4242)
    n = 8000
    rinv = pow(MONT, -1, p)
    a, b, c, d = ([_f2(rng, p) for _ in range(n)] for _ in range(4))
    a[0], b[0], c[0], d[0] = (p - 1, p - 1), (p - 1, p - 1), (p - 1, 0), (p - 1, p - 1)      # extremes of every partial sum
    a[1], b[1], c[1], d[1] = (p - 1, p - 1), (p - 1, 0), (0, p - 1), (p - 1, p - 1)
    A, B, Cc, D = _pack_f2(a), _pack_f2(b), _pack_f2(c), _pack_f2(d)
    out = np.zeros_like(A)

    def mul(x, y):      # Montgomery residues: (x y) / R
        return ((x[0] * y[0] - x[1] * y[1]) * rinv % p, (x[0] * y[1] + x[1] * y[0]) * rinv % p)

    want_mul = [mul(x, y) for x, y in zip(a, b)]
    want_msub = [tuple((u - v) % p for u, v in zip(mul(x, y), mul(z, w))) for x, y, z, w in zip(a, b, c, d)]
    for op, want in ((0, want_mul), (1, want_mul), (2, want_msub), (3, want_msub), (4, [mul(x, x) for x in a])):
        lib.fq2_host_op(op, ptr(A), ptr(B), ptr(Cc), ptr(D), ptr(out), n)
        assert lib.fp_host_carry_lost() == 0
        got = _unpack_f2(out)
        bad = [i for i in range(n) if got[i] != want[i]]
        assert not bad, "Fq2 op %d: first mismatch at %d" % (op, bad[0])


def _xyzz_of(curve_pts, rng, f_mul, f_sqr, scale):
    """affine oracle points -> XYZZ with a random non-trivial (zz, zzz) = (l^2, l^3), coordinates scaled accordingly"""
    out = []
    for pt, lam in zip(curve_pts, scale):
        l2 = f_sqr(lam)
        l3 = f_mul(l2, lam)
        out.append((f_mul(pt[0], l2), f_mul(pt[1], l3), l2, l3))
    return out


@pytest.mark.parametrize("group", [1, 2])
def test_madd_lazy_is_bit_identical_and_correct(lib, group):
    """acc + P through madd() and madd_lazy(): same words out, equal to the oracle's affine sum; the exceptional branches
    (identity accumulator, P == Q -> doubling, P == -Q -> identity) go through both forms as well."""
    p = bn.Q
    rng = random.Random(900 + group)
    n = 300
    fb = bn.fixed_base(group)
    ks = [rng.randrange(1, bn.R) for _ in range(n)]
    js = [rng.randrange(1, bn.R) for _ in range(n)]
    js[0], js[1] = ks[0], (bn.R - ks[1]) % bn.R                       # P == Q, P == -Q
    P = fb.mul_many(ks)
    Qp = fb.mul_many(js)
    if group == 1:
        f_mul = lambda a, b: a * b % p
        f_sqr = lambda a: a * a % p
        lam = [rng.randrange(1, p) for _ in range(n)]
        flat = lambda v: [v]
        W = 8
    else:
        f_mul, f_sqr = bn.f2_mul, bn.f2_sqr
        lam = [(rng.randrange(1, p), rng.randrange(p)) for _ in range(n)]
        flat = lambda v: [v[0], v[1]]
        W = 16
    acc = _xyzz_of(P, rng, f_mul, f_sqr, lam)
    zero = 0 if group == 1 else (0, 0)
    acc[2] = (zero, zero, zero, zero)                                # identity accumulator
    ACC = pack([c * MONT % p for a in acc for comp in a for c in flat(comp)])
    PT = pack_g1(Qp) if group == 1 else pack_g2(Qp)
    o0 = np.zeros(n * 4 * W * 4, dtype=np.uint8)
    o1 = np.zeros_like(o0)
    lib.xyzz_host_madd(group, 0, ptr(ACC), ptr(PT), ptr(o0), n)
    lib.xyzz_host_madd(group, 1, ptr(ACC), ptr(PT), ptr(o1), n)
    assert lib.fp_host_carry_lost() == 0
    assert o0.tobytes() == o1.tobytes(), "madd_lazy differs from madd"
    o2 = np.zeros_like(o0)
    lib.xyzz_host_madd(group, 2, ptr(ACC), ptr(PT), ptr(o2), n)
    assert lib.fp_host_carry_lost() == 0
    assert o0.tobytes() == o2.tobytes(), "madd_lazy with the dedicated squaring differs from madd"
    # against the oracle: x = X / ZZ, y = Y / ZZZ
    rinv = pow(MONT, -1, p)
    vals = [v * rinv % p for v in unpack(o1)]
    cv = fb.c
    for i in range(n):
        w = vals[i * 4 * (W // 8):(i + 1) * 4 * (W // 8)]
        if group == 1:
            X, Y, ZZ, ZZZ = w
            got = None if ZZ == 0 else (X * pow(ZZ, -1, p) % p, Y * pow(ZZZ, -1, p) % p)
        else:
            X, Y, ZZ, ZZZ = (w[0], w[1]), (w[2], w[3]), (w[4], w[5]), (w[6], w[7])
            got = None if ZZ == (0, 0) else (bn.f2_mul(X, bn.f2_inv(ZZ)), bn.f2_mul(Y, bn.f2_inv(ZZZ)))
        want = Qp[i] if i == 2 else cv.add(P[i], Qp[i])
        assert got == want, "group %d case %d" % (group, i)


@pytest.mark.parametrize("group", [1, 2])
def test_full_add_both_builds_identical_and_correct(lib, lib_tail, group):
    """XYZZ::add (bucket gather / reduction, blinding, tails) compiled with and without ZKR_LAZY_TAIL: same words out, equal
    to the oracle's sum; P == Q, P == -Q and identity operands included."""
    assert lib.fp_host_lazy_tail() == 0 and lib_tail.fp_host_lazy_tail() == 1
    p = bn.Q
    rng = random.Random(1700 + group)
    n = 200
    fb = bn.fixed_base(group)
    ks = [rng.randrange(1, bn.R) for _ in range(n)]
    js = [rng.randrange(1, bn.R) for _ in range(n)]
    js[0], js[1] = ks[0], (bn.R - ks[1]) % bn.R
    P, Qp = fb.mul_many(ks), fb.mul_many(js)
    if group == 1:
        f_mul, f_sqr = (lambda a, b: a * b % p), (lambda a: a * a % p)
        rnd = lambda: rng.randrange(1, p)
        flat = lambda v: [v]
        W = 8
    else:
        f_mul, f_sqr = bn.f2_mul, bn.f2_sqr
        rnd = lambda: (rng.randrange(1, p), rng.randrange(p))
        flat = lambda v: [v[0], v[1]]
        W = 16
    a = _xyzz_of(P, rng, f_mul, f_sqr, [rnd() for _ in range(n)])
    b = _xyzz_of(Qp, rng, f_mul, f_sqr, [rnd() for _ in range(n)])
    zero = 0 if group == 1 else (0, 0)
    a[2] = (zero,) * 4
    b[3] = (zero,) * 4
    A = pack([c * MONT % p for v in a for comp in v for c in flat(comp)])
    B = pack([c * MONT % p for v in b for comp in v for c in flat(comp)])
    o0 = np.zeros(n * 4 * W * 4, dtype=np.uint8)
    o1 = np.zeros_like(o0)
    lib.xyzz_host_add(group, ptr(A), ptr(B), ptr(o0), n)
    lib_tail.xyzz_host_add(group, ptr(A), ptr(B), ptr(o1), n)
    assert lib.fp_host_carry_lost() == 0 and lib_tail.fp_host_carry_lost() == 0
    assert o0.tobytes() == o1.tobytes(), "ZKR_LAZY_TAIL changes the words of XYZZ::add"
    cv = fb.c

    def affine(arr):
        rinv = pow(MONT, -1, p)
        vals = [v * rinv % p for v in unpack(arr)]
        res = []
        for i in range(n):
            w = vals[i * 4 * (W // 8):(i + 1) * 4 * (W // 8)]
            if group == 1:
                X, Y, ZZ, ZZZ = w
                res.append(None if ZZ == 0 else (X * pow(ZZ, -1, p) % p, Y * pow(ZZZ, -1, p) % p))
            else:
                X, Y, ZZ, ZZZ = (w[0], w[1]), (w[2], w[3]), (w[4], w[5]), (w[6], w[7])
                res.append(None if ZZ == (0, 0) else (bn.f2_mul(X, bn.f2_inv(ZZ)), bn.f2_mul(Y, bn.f2_inv(ZZZ))))
        return res

    got = affine(o1)
    for i in range(n):
        want = Qp[i] if i == 2 else P[i] if i == 3 else cv.add(P[i], Qp[i])
        assert got[i] == want, "group %d case %d" % (group, i)
    # doubling, both builds
    lib.xyzz_host_dbl(group, ptr(A), ptr(o0), n)
    lib_tail.xyzz_host_dbl(group, ptr(A), ptr(o1), n)
    assert lib.fp_host_carry_lost() == 0 and lib_tail.fp_host_carry_lost() == 0
    assert o0.tobytes() == o1.tobytes(), "ZKR_LAZY_TAIL changes the words of XYZZ::dbl"
    got = affine(o1)
    for i in range(n):
        assert got[i] == (None if i == 2 else cv.add(P[i], P[i])), "dbl, group %d case %d" % (group, i)
