"""Published BN254 known answers (EIP-196 / EIP-197 precompile test vectors, tests/golden/eip196_197_vectors.json) against
every implementation of the curve and pairing arithmetic in this repository: the Python-int oracle (oracle/bn254.py),
its C restatement (oracle/c), the library's host pairing (zkr_pairing_check, behind zkr_verify) and -- on a GPU -- the
device curve arithmetic (zkr_test_curve_op).  These vectors were not made by this repository: they remove the
self-reference on curve and pairing arithmetic.  The reference anchors are the precompile calls
of contracts/contracts/TxVerifier.sol:59-116 (staticcall 6 = ecAdd, 7 = ecMul, 8 = ecPairing)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import cbind
from simple_zk_rollups_b200 import _lib
from helpers import pack, pack_g1, unpack_g1

Q, R = bn.Q, bn.R
V = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "eip196_197_vectors.json")))


def H(s):
    return int(s, 16)


def g1(words):
    p = (H(words[0]), H(words[1]))
    assert (p[1] * p[1] - p[0] ** 3 - 3) % Q == 0, "transcription error: point off the curve"
    return p


def g2(words):
    """EIP-197 order: (x.imag, x.real, y.imag, y.real) -> ((c0, c1), (c0, c1)) with c0 the real part."""
    p = ((H(words[1]), H(words[0])), (H(words[3]), H(words[2])))
    x, y = p
    assert bn.f2_sqr(y) == bn.f2_add(bn.f2_mul(bn.f2_sqr(x), x), bn.B2), "transcription error: G2 point off the twist"
    assert bn.g2_in_subgroup(p)
    return p


def test_published_constants():
    c = V["constants"]
    assert H(c["q"]) == Q and H(c["r"]) == R
    assert bn.B2 == tuple(int(x) for x in c["twist_b_decimal"])
    assert g2(c["g2_generator"]) == bn.G2_GEN


def test_python_oracle_matches_published_vectors():
    fb = bn.fixed_base(1)
    for k, w in V["g1_multiples"].items():
        if k.isdigit():
            assert fb.mul_many([int(k)])[0] == g1(w)
            assert bn.G1.mul(bn.G1_GEN, int(k)) == g1(w)
    assert bn.G1.mul(bn.G1_GEN, R) is None
    for t in V["ecadd"]:
        assert bn.G1.add(g1(t["p"]), g1(t["q"])) == g1(t["sum"])
    for t in V["ecmul"]:
        assert bn.G1.mul(g1(t["p"]), H(t["k"])) == g1(t["product"])
    for t in V["ecpairing"]:
        w = t["input"]
        pairs = [(g1(w[i:i + 2]), g2(w[i + 2:i + 6])) for i in range(0, len(w), 6)]
        assert bn.pairing_product_is_one(pairs) is t["expected"]
        bad = [(bn.G1.add(pairs[0][0], bn.G1_GEN), pairs[0][1])] + pairs[1:]
        assert not bn.pairing_product_is_one(bad)


def test_c_oracle_matches_published_vectors():
    enc = lambda p: pack_g1([p]).tobytes()
    dec = lambda b: unpack_g1(np.frombuffer(b, dtype=np.uint8))[0]
    for t in V["ecadd"]:
        assert dec(cbind.g1_add(enc(g1(t["p"])), enc(g1(t["q"])))) == g1(t["sum"])
    for t in V["ecmul"]:
        assert dec(cbind.g1_mul(enc(g1(t["p"])), H(t["k"]))) == g1(t["product"])
    for k, w in V["g1_multiples"].items():
        if k.isdigit():
            assert dec(cbind.g1_mul(enc(bn.G1_GEN), int(k))) == g1(w)
    assert dec(cbind.g1_mul(enc(bn.G1_GEN), R)) is None


def _pairing_check(pairs):
    L = _lib.lib()
    a = pack([c for p, _ in pairs for c in p])
    b = pack([c for _, q in pairs for c in (q[0][0], q[0][1], q[1][0], q[1][1])])
    one = C.c_int(-1)
    _lib.check(L.zkr_pairing_check(_lib.buf_ptr(a), _lib.buf_ptr(b), len(pairs), C.byref(one)))
    return bool(one.value)


def test_library_host_pairing_matches_published_vectors():
    """zkr_pairing_check is host code (no GPU): the 4 x 64-bit-limb pairing behind zkr_verify."""
    for t in V["ecpairing"]:
        w = t["input"]
        pairs = [(g1(w[i:i + 2]), g2(w[i + 2:i + 6])) for i in range(0, len(w), 6)]
        assert _pairing_check(pairs) is t["expected"]
        assert not _pairing_check([(bn.G1.add(pairs[0][0], bn.G1_GEN), pairs[0][1])] + pairs[1:])
    # the EIP-197 'two_point_match' shape: e(P, Q) e(-P, Q) == 1 on published multiples of G1
    for k in ("2", "3", "9"):
        p = g1(V["g1_multiples"][k])
        assert _pairing_check([(p, bn.G2_GEN), (bn.G1.neg(p), bn.G2_GEN)])
        assert not _pairing_check([(p, bn.G2_GEN), (p, bn.G2_GEN)])


@pytest.mark.gpu
def test_device_curve_arithmetic_matches_published_vectors(zctx):
    L = _lib.lib()

    def op(code, ps, qs_or_ks, scalars=False):
        a = pack_g1(ps)
        b = pack(qs_or_ks) if scalars else pack_g1(qs_or_ks)
        out = np.zeros(64 * len(ps), dtype=np.uint8)
        _lib.check(L.zkr_test_curve_op(zctx, 1, code, _lib.buf_ptr(a), _lib.buf_ptr(b), _lib.buf_ptr(out), len(ps)))
        return unpack_g1(out)

    adds = V["ecadd"]
    assert op(0, [g1(t["p"]) for t in adds], [g1(t["q"]) for t in adds]) == [g1(t["sum"]) for t in adds]
    dbl = lambda p: bn.G1.add(p, p)
    assert op(3, [g1(t["p"]) for t in adds], [g1(t["q"]) for t in adds]) == \
        [bn.G1.add(dbl(g1(t["p"])), dbl(g1(t["q"]))) for t in adds]           # op 3 = 2P + 2Q through the full XYZZ add
    muls = V["ecmul"]
    assert op(2, [g1(t["p"]) for t in muls], [H(t["k"]) for t in muls], scalars=True) == [g1(t["product"]) for t in muls]
    ks = [int(k) for k in V["g1_multiples"] if k.isdigit()]
    assert op(2, [bn.G1_GEN] * len(ks), ks, scalars=True) == [g1(V["g1_multiples"][str(k)]) for k in ks]
    assert op(1, [bn.G1_GEN], [bn.G1_GEN]) == [g1(V["g1_multiples"]["2"])]
    assert op(0, [bn.G1_GEN], [g1(V["g1_multiples"]["2"])]) == [g1(V["g1_multiples"]["3"])]
