"""GPU parity: NTT / coset NTT / H pipeline over Fr vs the recursive radix-2 oracle, bit-exact."""
import ctypes as C
import random

import numpy as np
import pytest

from oracle import groth16 as g
from oracle.bn254 import R
from simple_zk_rollups_b200 import _lib
from helpers import pack, unpack

pytestmark = pytest.mark.gpu

FWD, INV, CFWD, CINV = 0, 1, 2, 3
BR_OUT, BR_IN = 0x10, 0x20


def brev(x, bits):
    return [x[g.bit_reverse(i, bits)] for i in range(len(x))]


def run(zctx, vals, log_n, mode):
    buf = pack(vals)
    _lib.check(_lib.lib().zkr_ntt(zctx, _lib.buf_ptr(buf), log_n, mode, 0))
    return unpack(buf)


@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 5, 6, 8, 10, 11, 12, 13, 14, 16])
def test_ntt_modes(zctx, log_n):
    rng = random.Random(log_n)
    n = 1 << log_n
    x = [rng.randrange(R) for _ in range(n)]
    x[0] = R - 1
    if n > 2:
        x[1] = 0
    shift = g.root_of_unity(log_n + 1)
    want = {FWD: g.ntt(x), INV: g.ntt(x, inverse=True), CFWD: g.coset_ntt(x, shift), CINV: g.coset_intt(x, shift)}
    for mode, w in want.items():
        assert run(zctx, x, log_n, mode) == w, "mode %d" % mode
        assert run(zctx, x, log_n, mode | BR_OUT) == brev(w, log_n), "mode %d bitrev out" % mode
        assert run(zctx, brev(x, log_n), log_n, mode | BR_IN) == w, "mode %d bitrev in" % mode


def test_ntt_roundtrip_large(zctx):
    """size-independent property at 2^20 / 2^22: iNTT(NTT(x)) == x and evaluation at one point."""
    L = _lib.lib()
    for log_n in (20, 22):
        n = 1 << log_n
        rs = np.random.RandomState(log_n)
        buf = rs.randint(0, 256, size=n * 32, dtype=np.uint8)
        buf.reshape(n, 32)[:, 31] &= 0x1F          # < 2^253 < r
        orig = buf.copy()
        _lib.check(L.zkr_ntt(zctx, _lib.buf_ptr(buf), log_n, FWD, 0))
        ev = buf.copy()
        _lib.check(L.zkr_ntt(zctx, _lib.buf_ptr(buf), log_n, INV, 0))
        assert np.array_equal(buf, orig)
        # Horner check of evaluation index k: X[k] = sum x_j w^(jk), for a sparse probe instead:
        # linearity probe -- NTT(e_5) has X[k] = w^(5k)
        probe = np.zeros(n * 32, dtype=np.uint8)
        probe[5 * 32] = 1
        _lib.check(L.zkr_ntt(zctx, _lib.buf_ptr(probe), log_n, FWD, 0))
        w = g.root_of_unity(log_n)
        for k in (0, 1, 2, 12345, n - 1):
            got = int.from_bytes(probe[k * 32:(k + 1) * 32].tobytes(), "little")
            assert got == pow(w, 5 * k, R)
        del ev


@pytest.mark.parametrize("log_m", [2, 3, 6, 10, 12, 13])
def test_h_pipeline(zctx, log_m):
    """h from A_T, B_T evaluations == oracle calc_h_lu / calc_h_websnark on the same vectors."""
    L = _lib.lib()
    rng = random.Random(100 + log_m)
    m = 1 << log_m
    at = [rng.randrange(R) for _ in range(m)]
    bt = [rng.randrange(R) for _ in range(m)]
    pk = dict(domainSize=m, polsA=[{i: at[i] for i in range(m)}], polsB=[{i: bt[i] for i in range(m)}])
    want = g.calc_h_lu(pk, [1])
    if log_m <= 10:
        assert want == g.calc_h_websnark(pk, [1])
    da, db, dh = C.c_void_p(), C.c_void_p(), C.c_void_p()
    for d in (da, db, dh):
        _lib.check(L.zkr_dev_malloc(zctx, m * 32, C.byref(d)))
    for bitrev in (0, 1):
        A, B = pack(at), pack(bt)
        _lib.check(L.zkr_dev_upload(zctx, da, _lib.buf_ptr(A), m * 32))
        _lib.check(L.zkr_dev_upload(zctx, db, _lib.buf_ptr(B), m * 32))
        _lib.check(L.zkr_h_from_evals_dev(zctx, da, db, log_m, dh, bitrev))
        out = np.zeros(m * 32, dtype=np.uint8)
        _lib.check(L.zkr_dev_download(zctx, _lib.buf_ptr(out), dh, m * 32))
        got = unpack(out)
        assert got == (brev(want, log_m) if bitrev else want)
    for d in (da, db, dh):
        _lib.check(L.zkr_dev_free(zctx, d))


@pytest.mark.parametrize("log_n", [20, 24])
def test_ntt_horner_at_random_points(zctx, log_n):
    """SURVEY 8(d) config 4 check at 2^20 / 2^24: uniform random input; the transformed values at 8 random indices
    (plus the first and the last) must equal the polynomial evaluated there by Horner's rule on the host (oracle C
    restatement, no transform code): forward X[k] = x(w^k), coset X[k] = x(g w^k), and the inverse transform's
    coefficients evaluated back at w^k must return the input."""
    from oracle import cbind
    L = _lib.lib()
    n = 1 << log_n
    rs = np.random.RandomState(100 + log_n)
    x = rs.randint(0, 256, size=n * 32, dtype=np.uint8)
    x.reshape(n, 32)[:, 31] &= 0x1F                          # < 2^253 < r
    w, gsh = g.root_of_unity(log_n), g.root_of_unity(log_n + 1)
    rng = random.Random(log_n)
    ks = [0, n - 1] + [rng.randrange(n) for _ in range(8)]
    val = lambda buf, k: int.from_bytes(buf[32 * k:32 * k + 32].tobytes(), "little")
    fwd = x.copy()
    _lib.check(L.zkr_ntt(zctx, _lib.buf_ptr(fwd), log_n, FWD, 0))
    for k in ks:
        assert val(fwd, k) == cbind.horner(x, pow(w, k, R)), "forward, index %d" % k
    del fwd
    cf = x.copy()
    _lib.check(L.zkr_ntt(zctx, _lib.buf_ptr(cf), log_n, CFWD, 0))
    for k in ks[:5]:
        assert val(cf, k) == cbind.horner(x, gsh * pow(w, k, R) % R), "coset forward, index %d" % k
    del cf
    inv = x.copy()
    _lib.check(L.zkr_ntt(zctx, _lib.buf_ptr(inv), log_n, INV, 0))
    for k in ks[:5]:
        assert cbind.horner(inv, pow(w, k, R)) == val(x, k), "inverse, index %d" % k


def test_fill_geometric_closed_form(zctx):
    """zkr_fill_geometric (the dense NTT workload whose transform bench.py checks in closed form):
    x_j = c g^j, and NTT(x)[k] = c (g^N - 1) / (g w^k - 1)."""
    L = _lib.lib()
    log_n = 14
    n = 1 << log_n
    c, gg = 0x1234567890ABCDEF1234567890ABCDEF % R, 0xFEDCBA9876543210FEDCBA987654321 % R
    d = C.c_void_p()
    _lib.check(L.zkr_dev_malloc(zctx, n * 32, C.byref(d)))
    cb, gb = pack([c]), pack([gg])                 # named: buf_ptr does not keep its argument alive
    _lib.check(L.zkr_fill_geometric(zctx, d, n, _lib.buf_ptr(cb), _lib.buf_ptr(gb), 0, log_n, 1, 0))
    out = np.zeros(n * 32, dtype=np.uint8)
    _lib.check(L.zkr_dev_download(zctx, _lib.buf_ptr(out), d, n * 32))
    assert unpack(out) == [c * pow(gg, j, R) % R for j in range(n)]
    _lib.check(L.zkr_ntt(zctx, d, log_n, FWD, 1))
    _lib.check(L.zkr_dev_download(zctx, _lib.buf_ptr(out), d, n * 32))
    w = g.root_of_unity(log_n)
    top = c * (pow(gg, n, R) - 1) % R
    got = unpack(out)
    for k in (0, 1, 77, n - 1):
        assert got[k] == top * pow((gg * pow(w, k, R) - 1) % R, -1, R) % R
    assert got == g.ntt([c * pow(gg, j, R) % R for j in range(n)])
    _lib.check(L.zkr_dev_free(zctx, d))
