"""GPU witness solver (zkr_wprog_build / zkr_witness_solve) against the host-side forward solve of
the synthetic rollup-shaped circuits (simple_zk_rollups_b200/synth.py: MiMC-Feistel rounds as in
prover/circuits/hasher.circom:8, products, constrained bits), value for value; and the witness-on-GPU proof generator
against the oracle's proof for the same (witness, r, s)."""
import ctypes as C
import random

import numpy as np
import pytest

from oracle import binfmt as bf
from oracle import groth16 as g
from oracle.bn254 import R
from simple_zk_rollups_b200 import _lib, keygen, prover, synth, witness

pytestmark = pytest.mark.gpu

TOXIC = (0x1234567890ABCDEF1234567890ABCDEF1234567, 0x2222222222222222222222222222222222221,
         0x3333333333333333333333333333333333333331, 0x44444444444444444444444444444444441,
         0x555555555555555555555555555555555555555551)


@pytest.fixture(scope="module")
def gp():
    p = prover.Groth16Prover(0)
    yield p
    p.close()


@pytest.mark.parametrize("nc,npub,seed", [(5, 1, 1), (40, 2, 2), (700, 5, 3), (5000, 17, 4)] + [synth.SHAPES["tx"] + (11,)])
def test_solver_reproduces_the_host_witness(gp, nc, npub, seed):
    r1, w = synth.generate(nc, npub, seed=seed)
    ws = witness.WitnessSolver(gp, r1)
    try:
        assert ws.n_vars == r1.nVars and ws.n_given + ws.n_solved == r1.nVars
        given = ws.given_signals.tolist()
        assert given[:npub + 4] == list(range(npub + 4))           # the constant 1, the public and the free inputs
        assert all(w[s] in (0, 1) for s in given[npub + 4:])        # everything else given is a constrained bit
        assert ws.solve(ws.given_from_witness(w)) == w
        # other inputs -> another satisfying witness of the same circuit
        rng = random.Random(seed)
        vals = ws.given_from_witness(w)
        vals[1:npub + 4] = [rng.randrange(R) for _ in range(npub + 3)]
        w2 = ws.solve(vals)
        assert w2 != w and w2[0] == 1
        if nc <= 5000:
            assert synth.check_witness(r1, w2)
        with pytest.raises(_lib.ZkrError) as ei:
            ws.solve([R] + vals[1:])
        assert ei.value.code == -3
    finally:
        ws.close()


def test_per_level_and_persistent_solves_agree(gp, monkeypatch):
    """ZKR_WITNESS_PER_LEVEL=1 (one launch per level) and the default persistent kernel (grid barrier between levels)."""
    r1, w = synth.generate(2500, 6, seed=21)
    ws = witness.WitnessSolver(gp, r1)
    try:
        given = ws.given_from_witness(w)
        assert ws.solve(given) == w
        monkeypatch.setenv("ZKR_WITNESS_PER_LEVEL", "1")
        assert ws.solve(given) == w
    finally:
        ws.close()


def test_same_circuit_different_witness_seed(gp):
    """synth.generate(witness_seed=k): same circuit, the solver must reproduce each witness from its given values."""
    r1, wa = synth.generate(900, 4, seed=5, witness_seed=1)
    _, wb = synth.generate(900, 4, seed=5, witness_seed=2)
    ws = witness.WitnessSolver(gp, r1)
    try:
        assert ws.solve(ws.given_from_witness(wa)) == wa
        assert ws.solve(ws.given_from_witness(wb)) == wb
    finally:
        ws.close()


def test_not_forward_solvable_is_rejected(gp):
    """A row that needs a signal only a later row defines is refused at build time, not mis-solved."""
    r1, w = synth.generate(12, 1, seed=7)
    sig, row, cid = r1.mats["C"]
    # swap the rows of the first two MiMC constraints in every matrix: t4 = t2 * t2 now comes before t2 is defined
    def swap(m):
        s, r_, c = m
        r2 = r_.copy()
        r2[r_ == 0] = 1
        r2[r_ == 1] = 0
        o = np.lexsort((r2, s))
        return s[o], r2[o], c[o]
    r1.mats = {k: swap(v) for k, v in r1.mats.items()}
    with pytest.raises(_lib.ZkrError) as ei:
        witness.WitnessSolver(gp, r1)
    assert ei.value.code == -8


def test_gpu_witness_proof_generator_matches_oracle(gp):
    r1, w = synth.generate(300, 4, seed=9)
    pk, vk, _ = g.setup(r1.to_dicts(), TOXIC)
    key = gp.load_key(bf.binarify_proving_key(pk))
    vkey = gp.load_vkey(bf.vk_to_json(vk))
    ws = witness.WitnessSolver(gp, r1)
    try:
        gen = witness.createGpuProofGenerator(gp, key, ws, 4, vkey=vkey)
        out = gen(ws.given_from_witness(w), r=1234567, s=7654321)
        want, pub = g.gen_proof(pk, w, 1234567, 7654321)
        assert out["proof_bytes"] == g.proof_to_bytes(want)
        assert out["publicSignals"] == [str(x) for x in pub]
        rnd = gen(ws.given_from_witness(w))                      # CSPRNG blinding: another valid proof of the same statement
        assert rnd["proof_bytes"] != out["proof_bytes"]
        assert g.verify(vk, g.proof_from_bytes(rnd["proof_bytes"]), pub)
    finally:
        ws.close()
