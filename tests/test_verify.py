"""Tests of the native verifier (csrc/verify.cu).

CPU part (`-m "not gpu"`): zkr_pairing_check, the host pairing product behind zkr_verify, against the oracle's
independent pairing (oracle/bn254.py: big-int Fq12, plain (q^12-1)/r exponent) and on the reference's own
committed constants -- generators (TxVerifier.sol:24-35) and the two verifying keys in the verifier
contracts (TxVerifier.sol:177-255, WithdrawVerifier.sol:177-185).
GPU part: zkr_verify (GPU vk_x MSM + host pairing product) = the acceptance predicate TxVerifier.sol:258-276,
with the tamper-negatives of contracts/__tests__/withdrawverifier.test.ts:42-68."""
import ctypes as C
import json
import os
import random

import numpy as np
import pytest

from oracle import binfmt as bf
from oracle import bn254 as bn
from oracle import groth16 as g
from simple_zk_rollups_b200 import _lib, binarify, keygen, prover, synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_constants.json")))
R, Q = bn.R, bn.Q
TOXIC = (1234567891011, 222222222222223, 3333333333333331, 44444444444447, 5555555555555557)


def _g2_sol(s):
    return ((int(s[0][1]), int(s[0][0])), (int(s[1][1]), int(s[1][0])))


def _b1(p):
    return b"\0" * 64 if p is None else int(p[0]).to_bytes(32, "little") + int(p[1]).to_bytes(32, "little")


def _b2(p):
    if p is None:
        return b"\0" * 128
    return b"".join(int(v).to_bytes(32, "little") for v in (p[0][0], p[0][1], p[1][0], p[1][1]))


def _f2_pow(a, e):
    r = bn.F2_ONE
    for bit in bin(e)[2:]:
        r = bn.f2_sqr(r)
        if bit == "1":
            r = bn.f2_mul(r, a)
    return r


def _f2_sqrt(a):
    """square root in Fq2 for q = 3 mod 4 (None if a is a non-residue)"""
    a1 = _f2_pow(a, (Q - 3) // 4)
    alpha = bn.f2_mul(bn.f2_sqr(a1), a)
    if bn.f2_mul(bn.f2_conj(alpha), alpha) == ((Q - 1) % Q, 0):
        return None
    x0 = bn.f2_mul(a1, a)
    if alpha == ((Q - 1) % Q, 0):
        x = bn.f2_mul((0, 1), x0)
    else:
        x = bn.f2_mul(_f2_pow(bn.f2_add(bn.F2_ONE, alpha), (Q - 1) // 2), x0)
    return x if bn.f2_sqr(x) == a else None


def pairing_check(pairs):
    L = _lib.lib()
    g1 = b"".join(_b1(p) for p, _ in pairs)
    g2 = b"".join(_b2(q) for _, q in pairs)
    ok = C.c_int(-1)
    rc = L.zkr_pairing_check(g1, g2, len(pairs), C.byref(ok))
    return rc, ok.value


def test_pairing_check_bilinearity_on_generators():
    G1, G2 = bn.G1_GEN, bn.G2_GEN
    a, b = 0xDEADBEEF12345, 0xC0FFEE987
    # e(aG1, bG2) * e(-abG1, G2) == 1
    assert pairing_check([(bn.G1.mul(G1, a), bn.G2.mul(G2, b)), (bn.G1.neg(bn.G1.mul(G1, a * b % R)), G2)]) == (0, 1)
    # and with the factor off by one it is not
    assert pairing_check([(bn.G1.mul(G1, a), bn.G2.mul(G2, b)), (bn.G1.neg(bn.G1.mul(G1, (a * b + 1) % R)), G2)]) == (0, 0)
    # non-degenerate: a single pairing of the generators is not 1
    assert pairing_check([(G1, G2)]) == (0, 0)
    # e(P, Q) e(-P, Q) == 1, infinity contributes 1, the empty product is 1
    assert pairing_check([(G1, G2), (bn.G1.neg(G1), G2)]) == (0, 1)
    assert pairing_check([(None, G2), (G1, None)]) == (0, 1)
    assert pairing_check([]) == (0, 1)


def test_pairing_check_on_the_reference_verifying_keys():
    """Bilinearity through the points the reference commits (IC / alfa1 in G1; beta2, gamma2, delta2 in G2)."""
    rng = random.Random(3)
    for which in ("tx_vk", "withdraw_vk"):
        vk = GOLD[which]
        g1s = [tuple(int(c) for c in p) for p in vk["IC"][:3]] + [tuple(int(c) for c in vk["alfa1"])]
        g2s = [_g2_sol(vk[n + "_sol_order"]) for n in ("beta2", "gamma2", "delta2")]
        for p in g1s[:2] + g1s[-1:]:
            for q2 in g2s:
                k = rng.randrange(1, R)
                assert pairing_check([(bn.G1.mul(p, k), q2), (bn.G1.neg(p), bn.G2.mul(q2, k))]) == (0, 1)
        # 4-pair product with known exponents: e(aP,Q) e(bP,Q) e(cP,Q) e(-(a+b+c)P,Q) == 1
        p, q2 = g1s[0], g2s[0]
        a, b, c = (rng.randrange(R) for _ in range(3))
        pairs = [(bn.G1.mul(p, a), q2), (bn.G1.mul(p, b), q2), (bn.G1.mul(p, c), q2),
                 (bn.G1.neg(bn.G1.mul(p, (a + b + c) % R)), q2)]
        assert pairing_check(pairs) == (0, 1)
        pairs[2] = (bn.G1.mul(p, (c + 1) % R), q2)
        assert pairing_check(pairs) == (0, 0)


def test_pairing_check_agrees_with_the_oracle_on_random_products():
    rng = random.Random(9)
    for trial in range(6):
        ks = [rng.randrange(1, R) for _ in range(3)]
        ls = [rng.randrange(1, R) for _ in range(3)]
        pairs = [(bn.G1.mul(bn.G1_GEN, k), bn.G2.mul(bn.G2_GEN, l)) for k, l in zip(ks, ls)]
        tot = sum(k * l for k, l in zip(ks, ls)) % R
        if trial % 2:
            tot = (tot + trial) % R
        pairs.append((bn.G1.neg(bn.G1.mul(bn.G1_GEN, tot)), bn.G2_GEN))
        want = bn.pairing_product_is_one(pairs)
        assert want == (trial % 2 == 0)
        assert pairing_check(pairs) == (0, int(want))


def test_pairing_variants_agree():
    """The default host pairing (projective Miller loop, complex / cyclotomic squarings) against the first
    implementation (affine loop with batched inversions, generic squarings: ZKR_PAIRING_AFFINE=1 ZKR_PAIRING_CYC=0) and
    against itself with the cyclotomic-squaring self-check on (ZKR_PAIRING_CYC=2 aborts on a mismatch).  The knobs are
    read once per process, so the variants run in child processes on the same byte strings."""
    import os
    import subprocess
    import sys
    rng = random.Random(17)
    cases = []
    for trial in range(8):
        n = 1 + trial % 4
        ks = [rng.randrange(1, R) for _ in range(n)]
        ls = [rng.randrange(1, R) for _ in range(n)]
        pairs = [(bn.G1.mul(bn.G1_GEN, k), bn.G2.mul(bn.G2_GEN, l)) for k, l in zip(ks, ls)]
        tot = sum(k * l for k, l in zip(ks, ls)) % R
        if trial % 3 == 1:
            tot = (tot + 1) % R                                # product != 1
        pairs.append((bn.G1.neg(bn.G1.mul(bn.G1_GEN, tot)), bn.G2_GEN))
        if trial == 5:
            pairs.insert(1, (None, bn.G2_GEN))                 # infinity contributes 1
        cases.append((pairs, trial % 3 != 1))
    blobs = [(b"".join(_b1(p) for p, _ in pairs).hex(), b"".join(_b2(q) for _, q in pairs).hex(), len(pairs)) for pairs, _ in cases]
    here = [pairing_check(pairs) for pairs, _ in cases]
    assert here == [(0, int(w)) for _, w in cases]
    child = ("import sys, json, ctypes as C\n"
             "sys.path.insert(0, %r)\n"
             "from simple_zk_rollups_b200 import _lib\n"
             "L = _lib.lib()\n"
             "out = []\n"
             "for g1, g2, n in json.loads(sys.stdin.read()):\n"
             "    ok = C.c_int(-1)\n"
             "    rc = L.zkr_pairing_check(bytes.fromhex(g1), bytes.fromhex(g2), n, C.byref(ok))\n"
             "    out.append([rc, ok.value])\n"
             "print(json.dumps(out))\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for env in ({"ZKR_PAIRING_AFFINE": "1", "ZKR_PAIRING_CYC": "0"}, {"ZKR_PAIRING_CYC": "2"}, {"ZKR_PAIRING_AFFINE": "1"}):
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, "-c", child], input=json.dumps(blobs), capture_output=True, text=True, env=e, timeout=300)
        assert r.returncode == 0, r.stderr[-500:]
        assert [tuple(x) for x in json.loads(r.stdout.strip().splitlines()[-1])] == here, env


def test_pairing_check_rejects_invalid_points():
    G1, G2 = bn.G1_GEN, bn.G2_GEN
    assert pairing_check([((1, 3), G2)])[0] == -1                       # not on the G1 curve
    assert pairing_check([((Q, 2), G2)])[0] == -1                       # coordinate >= q
    assert pairing_check([(G1, ((1, 0), (2, 0)))])[0] == -1             # not on the twist
    # on the twist but outside the order-r subgroup (the twist has a large cofactor): first x = 1, 2, .. with a root
    pt = None
    for x in range(1, 60):
        rhs = bn.f2_add(bn.f2_mul(bn.f2_sqr((x, 0)), (x, 0)), bn.B2)
        y = _f2_sqrt(rhs)
        if y is not None:
            pt = ((x, 0), y)
            break
    assert pt is not None and bn.G2.is_on_curve(pt) and not bn.g2_in_subgroup(pt)
    assert pairing_check([(G1, pt)])[0] == -1
    L = _lib.lib()
    assert L.zkr_pairing_check(None, None, 9, C.byref(C.c_int())) == -1


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def gp():
    p = prover.Groth16Prover(0)
    yield p
    p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nc,npub,seed", [(30, 3, 1), (200, 10, 3), (64, 0, 5)])
def test_verify_accepts_honest_rejects_tampered(gp, nc, npub, seed):
    r1, w = synth.generate(nc, npub, seed=seed)
    pk, vk, _ = g.setup(r1.to_dicts(), TOXIC)
    key = gp.load_key(bf.binarify_proving_key(pk))
    vkey = gp.load_vkey(json.dumps(bf.vk_to_json(vk)))
    proof, _ = gp.prove(key, bf.binarify_witness(w), 77, 99)
    pub = w[1:npub + 1]
    assert g.verify(vk, g.proof_from_bytes(proof), pub)
    assert gp.verify(vkey, proof, pub) is True
    assert gp.verify(vkey, binarify.proof_from_bytes(proof), [str(x) for x in pub]) is True     # object / string forms
    # tampered public input (withdrawverifier.test.ts:42-68)
    for i in range(min(npub, 3)):
        bad = list(pub)
        bad[i] = (bad[i] + 1) % R
        assert gp.verify(vkey, proof, bad) is False
    # tampered proof points: swap A and C, perturb a coordinate (off curve), another valid point
    pb = bytearray(proof)
    swapped = bytes(pb[192:256] + pb[64:192] + pb[0:64])
    assert gp.verify(vkey, swapped, pub) is False
    off = bytearray(proof)
    off[0] ^= 1
    assert gp.verify(vkey, bytes(off), pub) is False
    other = bytearray(proof)
    other[192:256] = _b1(bn.G1.mul(bn.G1_GEN, 12345))
    assert gp.verify(vkey, bytes(other), pub) is False
    big = bytearray(proof)
    big[0:32] = (Q + 1).to_bytes(32, "little")
    assert gp.verify(vkey, bytes(big), pub) is False
    # contract reverts: wrong input count, input >= r
    with pytest.raises(_lib.ZkrError) as e:
        gp.verify(vkey, proof, list(pub) + [1])
    assert e.value.code == -1 and "verifier-bad-input" in str(e.value)
    if npub:
        with pytest.raises(_lib.ZkrError) as e:
            gp.verify(vkey, proof, [R] + list(pub[1:]))
        assert e.value.code == -3


@pytest.mark.gpu
def test_verify_binary_vkey_and_generator_self_check(gp):
    """vk block of the GPU setup -> zkr_vkey_load_bin; createProofGenerator's default isValid is zkr_verify and
    raises 'Invalid proof generated' (common.ts:36-38) when the witness violates the circuit."""
    r1, w = synth.generate(120, 5, seed=6)
    pk_bin, vk = keygen.synth_setup(gp.ctx, r1, TOXIC)
    key = gp.load_key(pk_bin)
    vkey = gp.load_vkey(vk["bin"])
    proof, _ = gp.prove(key, synth.witness_bytes(w), 5, 6)
    assert gp.verify(vkey, proof, w[1:6])
    assert g.verify(vk, g.proof_from_bytes(proof), w[1:6])
    vkj = bf.vk_to_json({k: v for k, v in vk.items() if k != "bin"})
    state = {"w": w}
    gen = prover.createProofGenerator(pk_bin, vkj, "tx.circom", lambda name, inp: (state["w"], 5), prover=gp)
    out = gen({})
    assert len(out["solidityProof"]["inputs"]) == 5
    w2 = list(w)
    w2[20] = (w2[20] + 1) % R
    state["w"] = w2
    with pytest.raises(RuntimeError, match="Invalid proof generated"):
        gen({})
    with pytest.raises(_lib.ZkrError) as e:
        gp.load_vkey(vk["bin"][:-1])
    assert e.value.code == -2
    bad = bytearray(vk["bin"])
    bad[5] ^= 0x40                                   # alfa1 off the curve
    with pytest.raises(_lib.ZkrError) as e:
        gp.load_vkey(bytes(bad))
    assert e.value.code == -2


@pytest.mark.gpu
def test_verify_full_size_public_inputs(gp):
    """tx.circom shape (73 public signals, BASELINE configs[0]): the vk_x MSM at the reference's real width."""
    nc, npub = synth.SHAPES["tx"]
    r1, w = synth.generate(nc, npub, seed=11)
    pk_bin, vk = keygen.synth_setup(gp.ctx, r1, TOXIC)
    key = gp.load_key(pk_bin)
    vkey = gp.load_vkey(vk["bin"])
    proof, _ = gp.prove(key, np.frombuffer(synth.witness_bytes(w), dtype=np.uint8), 123, 456)
    pub = w[1:npub + 1]
    assert gp.verify(vkey, proof, pub)
    bad = list(pub)
    bad[-1] = (bad[-1] + 1) % R
    assert not gp.verify(vkey, proof, bad)
    gp.L.zkr_pkey_free(key)
    gp._keys.remove(key)
