"""CPU tests: pin the oracle against every constant / key the reference holds for this path and against
itself (independent algorithms must agree).  PARITY UNPINNED upstream -- see DESIGN.md."""
import json
import os
import random

import pytest

from oracle import binfmt as bf
from oracle import bn254 as bn
from oracle import groth16 as g
from simple_zk_rollups_b200 import binarify, synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_constants.json")))
R, Q = bn.R, bn.Q
TOXIC = (11, 22, 33, 44, 55)


def _g2_sol(s):          # Solidity lists each Fq2 coordinate imaginary part first (TxVerifier.sol:18)
    return ((int(s[0][1]), int(s[0][0])), (int(s[1][1]), int(s[1][0])))


def test_reference_constants():
    assert int(GOLD["q"]) == Q == int(GOLD["q_sol"])
    assert int(GOLD["r"]) == R == int(GOLD["r_sol"]) == int(GOLD["snark_field_size_crypto_ts"])
    assert tuple(int(x) for x in GOLD["g1_generator"]) == bn.G1_GEN
    assert _g2_sol(GOLD["g2_generator_sol_order"]) == bn.G2_GEN
    assert GOLD["tx_n_public"] == 73 == synth.rollup_shape(GOLD["tx_batch_size"], GOLD["tx_tree_depth"])[1]
    assert len(GOLD["tx_vk"]["IC"]) == 74 and len(GOLD["withdraw_vk"]["IC"]) == 4
    # SURVEY.md B.1 known answers
    assert g.root_of_unity(28) == 19103219067921713944291392827692070036145651957329286315305642004821462161904
    assert g.root_of_unity(17) == 12650941915662020058015862023665998998969191525479888727406889100124684769509
    assert g.root_of_unity(20) == 17220337697351015657950521176323262483320249231368149235373741788599650842711
    assert (1 << 256) % Q == 6350874878119819312338956282401532409788428879151445726012394534686998597021
    assert (1 << 256) % R == 6350874878119819312338956282401532410528162663560392320966563075034087161851


@pytest.mark.parametrize("which", ["tx_vk", "withdraw_vk"])
def test_committed_verifying_keys_are_valid_group_elements(which):
    """The two verifying keys committed in the reference's verifier contracts: every point on curve and in
    the order-r subgroup (fixtures for the oracle's field / curve arithmetic)."""
    vk = GOLD[which]
    pts = [tuple(int(c) for c in p) for p in vk["IC"]] + [tuple(int(c) for c in vk["alfa1"])]
    for p in pts:
        assert bn.G1.is_on_curve(p)
    for p in pts[:6]:
        assert bn.g1_in_subgroup(p)
    for n in ("beta2", "gamma2", "delta2"):
        assert bn.g2_in_subgroup(_g2_sol(vk[n + "_sol_order"]))


def test_pairing_bilinear_and_nondegenerate():
    e = bn.pairing(bn.G1_GEN, bn.G2_GEN)
    assert e != bn.f12_one() and bn.f12_pow(e, R) == bn.f12_one()
    a, b = 0xDEADBEEF12345, 0xC0FFEE987
    assert bn.pairing(bn.G1.mul(bn.G1_GEN, a), bn.G2.mul(bn.G2_GEN, b)) == bn.f12_pow(e, a * b % R)
    vk = GOLD["tx_vk"]
    ic1 = tuple(int(c) for c in vk["IC"][1])
    b2 = _g2_sol(vk["beta2_sol_order"])
    assert bn.pairing_product_is_one([(bn.G1.mul(ic1, 5), b2), (bn.G1.neg(ic1), bn.G2.mul(b2, 5))])


def test_ntt_against_direct_dft():
    rng = random.Random(1)
    for bits in (1, 2, 3, 5):
        n = 1 << bits
        x = [rng.randrange(R) for _ in range(n)]
        w = g.root_of_unity(bits)
        dft = [sum(x[j] * pow(w, j * k, R) for j in range(n)) % R for k in range(n)]
        assert g.ntt(x) == dft
        assert g.ntt(dft, inverse=True) == x
        sh = g.root_of_unity(bits + 1)
        assert g.coset_intt(g.coset_ntt(x, sh), sh) == x


@pytest.mark.parametrize("nc,npub", [(5, 1), (40, 3), (130, 6)])
def test_h_methods_agree_and_proof_verifies(nc, npub):
    r1, w = synth.generate(nc, npub, seed=nc)
    assert synth.check_witness(r1, w)
    pk, vk, sec = g.setup(r1.to_dicts(), TOXIC)
    h, low = g.calc_h_snarkjs(pk, w)
    assert all((x + y) % R == 0 for x, y in zip(h, low)) and h[-1] == 0
    assert h == g.calc_h_websnark(pk, w) == g.calc_h_coset(pk, w) == g.calc_h_lu(pk, w)
    rng = random.Random(nc)
    r, s = rng.randrange(R), rng.randrange(R)
    proof, pub = g.gen_proof(pk, w, r, s)
    assert g.verify(vk, proof, pub)
    assert g.exponent_check(pk, sec, w, proof, r, s)
    assert g.exponent_check_flat(r1.with_input_rows(), r1.pool, npub, w, TOXIC, pk["domainSize"], proof, r, s)
    bad = list(pub)
    bad[0] = (bad[0] + 1) % R
    assert not g.verify(vk, proof, bad)                      # withdrawverifier.test.ts:42-68
    assert not g.verify(vk, proof, pub + [1])
    assert not g.verify(vk, proof, [R] + pub[1:])            # TxVerifier.sol:265
    # different (r, s) -> different but equally valid proof
    p2, _ = g.gen_proof(pk, w, 1, 2)
    assert p2 != proof and g.verify(vk, p2, pub)


def test_wire_formats_roundtrip_and_match_host_mirror():
    r1, w = synth.generate(60, 3, seed=3)
    pk, vk, _ = g.setup(r1.to_dicts(), TOXIC)
    blob = bf.binarify_proving_key(pk)
    assert blob == binarify.binarifyProvingKey(bf.pk_to_json(pk))
    assert bf.binarify_witness(w) == binarify.binarifyWitness([str(x) for x in w]) == synth.witness_bytes(w)
    back = bf.parse_proving_key(blob)
    for k in ("polsA", "polsB", "A", "B1", "B2", "C", "hExps", "vk_alfa_1", "vk_delta_2"):
        assert back[k] == pk[k], k
    assert any(p is None for p in pk["B1"])                  # infinity bases exist and survive the round trip
    n, l, m = pk["nVars"], pk["nPublic"], pk["domainSize"]
    nnz = sum(len(c) for c in pk["polsA"]) + sum(len(c) for c in pk["polsB"])
    assert len(blob) == 40 + 192 + 256 + 36 * nnz + 8 * n + 64 * n * 2 + 128 * n + 64 * (n - l - 1) + 64 * m
    proof, _ = g.gen_proof(pk, w, 9, 8)
    assert g.proof_from_bytes(g.proof_to_bytes(proof)) == proof
    assert bf.proof_from_json(bf.proof_to_json(proof)) == proof
    assert bf.proof_from_json(binarify.proof_from_bytes(g.proof_to_bytes(proof))) == proof
    assert bf.vk_from_json(bf.vk_to_json(vk)) == vk


def test_synth_generator_shapes():
    assert synth.SHAPES["tx"] == (107300, 73)
    assert synth.SHAPES["tx_2p20"] == (858400, 577)
    r1, w = synth.generate(2763, 3, seed=0)
    assert r1.domain() == (12, 4096) and r1.nVars == 1 + 3 + 3 + 2763
    assert synth.check_witness(r1, w)
    bits01 = sum(1 for x in w if x in (0, 1))
    assert 0.02 < bits01 / len(w) < 0.05                      # ~3 % {0,1} signals
    r2, w2 = synth.generate(2763, 3, seed=0)
    assert w == w2 and all((a == b).all() for k in "ABC" for a, b in zip(r1.mats[k], r2.mats[k]))
