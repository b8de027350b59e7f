"""Known-answer vectors of the whole hot path (tests/golden/groth16_kats.json, made by tools/make_golden_kats.py
from the oracle's pure-Python leg; the reference commits no prover vectors -- SURVEY.md 8(c)).

CPU: the C port reproduces every vector; the Python oracle reproduces itself (drift guard); the product's host
pairing accepts the honest proofs and rejects the circuit-violating one.
GPU: zkr_prove reproduces every vector from the committed binary keys, zkr_verify agrees, and the GPU setup
regenerates the committed keys byte for byte (sha256)."""
import base64
import ctypes as C
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import binfmt as bf
from oracle import bn254 as bn
from oracle import cbind
from oracle import groth16 as g
from simple_zk_rollups_b200 import _lib, keygen, prover, synth

KATS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "groth16_kats.json")))
TOXIC = tuple(int(t) for t in KATS["toxic"])
VECS = KATS["vectors"]
EMBEDDED = [v for v in VECS if "pk_bin_gz_b64" in v]


def _pk_bin(v):
    b = gzip.decompress(base64.b64decode(v["pk_bin_gz_b64"]))
    assert hashlib.sha256(b).hexdigest() == v["pk_bin_sha256"] and len(b) == v["pk_bin_len"]
    return b


def _witness(v, invalid=False):
    w = [int(x) for x in v["witness"]]
    if invalid:
        w[v["invalid_witness"]["index"]] = int(v["invalid_witness"]["witness_value"])
    return w


@pytest.mark.parametrize("v", EMBEDDED, ids=lambda v: "m%d" % v["domain_size"])
def test_c_port_reproduces_kats(v):
    pk_bin, w = _pk_bin(v), _witness(v)
    wb = bf.binarify_witness(w)
    for p in v["proofs"]:
        for mode in (0, 1):           # snarkjs arithmetic structure / Pippenger + iterative NTT
            got = cbind.prove(pk_bin, wb, int(p["r"]), int(p["s"]), mode=mode, threads=2)
            assert got.hex() == p["proof_hex"], (v["domain_size"], p["r"][:8], mode)
    iv = v["invalid_witness"]
    got = cbind.prove(pk_bin, bf.binarify_witness(_witness(v, True)), int(iv["r"]), int(iv["s"]), mode=1, threads=2)
    assert got.hex() == iv["proof_hex"]


def test_python_oracle_reproduces_smallest_kats():
    for v in EMBEDDED[:2]:
        pk = bf.parse_proving_key(_pk_bin(v))
        w = _witness(v)
        for p in v["proofs"]:
            proof, _ = g.gen_proof(pk, w, int(p["r"]), int(p["s"]), h_method=g.calc_h_websnark)   # parsed keys carry no polsC
            assert g.proof_to_bytes(proof).hex() == p["proof_hex"]


def test_keys_regenerate_from_the_oracle_setup():
    v = VECS[0]
    r1, w = synth.generate(v["n_constraints"], v["n_public"], seed=v["seed"])
    assert [str(x) for x in w] == v["witness"]
    pk, vk, _ = g.setup(r1.to_dicts(), TOXIC)
    assert hashlib.sha256(bf.binarify_proving_key(pk)).hexdigest() == v["pk_bin_sha256"]
    assert bf.vk_to_json(vk) == v["vk"]


@pytest.mark.parametrize("v", VECS[:3], ids=lambda v: "m%d" % v["domain_size"])
def test_host_pairing_accepts_kats_and_rejects_the_invalid_one(v):
    """TxVerifier.sol:258-276 with vk_x from the oracle's G1 arithmetic and the product's zkr_pairing_check."""
    L = _lib.lib()
    vk = bf.vk_from_json(v["vk"])
    pub = [int(x) for x in v["witness"][1:v["n_public"] + 1]]
    vkx = bn.G1.to_jac(vk["IC"][0])
    for x, ic in zip(pub, vk["IC"][1:]):
        vkx = bn.G1.jadd(vkx, bn.G1.jmul(bn.G1.to_jac(ic), x))
    vkx = bn.G1.to_affine(vkx)

    def b1(p):
        return int(p[0]).to_bytes(32, "little") + int(p[1]).to_bytes(32, "little")

    def b2(p):
        return b"".join(int(c).to_bytes(32, "little") for c in (p[0][0], p[0][1], p[1][0], p[1][1]))

    def check(proof_hex):
        pr = g.proof_from_bytes(bytes.fromhex(proof_hex))
        g1 = b1(bn.G1.neg(pr["pi_a"])) + b1(vk["vk_alfa_1"]) + b1(vkx) + b1(pr["pi_c"])
        g2 = b2(pr["pi_b"]) + b2(vk["vk_beta_2"]) + b2(vk["vk_gamma_2"]) + b2(vk["vk_delta_2"])
        ok = C.c_int(-1)
        assert L.zkr_pairing_check(g1, g2, 4, C.byref(ok)) == 0
        return ok.value

    for p in v["proofs"]:
        assert check(p["proof_hex"]) == 1
    assert check(v["invalid_witness"]["proof_hex"]) == 0


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def gp():
    p = prover.Groth16Prover(0)
    yield p
    p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("v", VECS, ids=lambda v: "m%d" % v["domain_size"])
def test_gpu_reproduces_kats(gp, v):
    r1, w = synth.generate(v["n_constraints"], v["n_public"], seed=v["seed"])
    assert [str(x) for x in w] == v["witness"]
    pk_gpu, vk_gpu = keygen.synth_setup(gp.ctx, r1, TOXIC)
    assert hashlib.sha256(pk_gpu.tobytes()).hexdigest() == v["pk_bin_sha256"], "GPU setup != committed key"
    pk_bin = _pk_bin(v) if "pk_bin_gz_b64" in v else pk_gpu.tobytes()
    key = gp.load_key(pk_bin)
    vkey = gp.load_vkey(json.dumps(v["vk"]))
    wb = bf.binarify_witness(w)
    pub = w[1:v["n_public"] + 1]
    for p in v["proofs"]:
        got, _ = gp.prove(key, wb, int(p["r"]), int(p["s"]))
        assert got.hex() == p["proof_hex"], (v["domain_size"], p["r"][:8])
        assert gp.verify(vkey, got, pub)
    iv = v["invalid_witness"]
    got, _ = gp.prove(key, bf.binarify_witness(_witness(v, True)), int(iv["r"]), int(iv["s"]))
    assert got.hex() == iv["proof_hex"]
    assert not gp.verify(vkey, got, pub)
