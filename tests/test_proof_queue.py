"""ProofQueue (simple_zk_rollups_b200/proof_queue.py): the multi-GPU prove queue an operator batch loop drains.
CPU: queue / ordering / error-delivery logic with stand-in provers.  GPU: real proofs over two contexts, each
self-checked with zkr_verify, equal to the oracle's."""
import random
import threading
import time

import pytest

from simple_zk_rollups_b200 import proof_queue

R = proof_queue.SNARK_FIELD_SIZE


class _FakeProver:
    def __init__(self, device, delay):
        self.device, self.delay, self.calls, self.lock = device, delay, 0, threading.Lock()

    def key_info(self, key):
        return {"nPublic": 2}

    def prove(self, key, wbin, r, s):
        time.sleep(self.delay)
        with self.lock:
            self.calls += 1
        w1 = int.from_bytes(bytes(wbin[32:64]), "little")
        if w1 == 999:
            raise ValueError("bad witness")
        vals = [w1, r, s, self.device, 0, 0, 0, 0]
        return b"".join(int(v).to_bytes(32, "little") for v in vals), {"total_ms": 1.0}

    def verify(self, vkey, proof, pub):
        return pub[0] != 13


def _w(x):
    return b"".join(int(v).to_bytes(32, "little") for v in (1, x, 7, 5))


def test_queue_logic_with_stand_in_provers():
    provers = [_FakeProver(0, 0.01), _FakeProver(1, 0.03)]
    q = proof_queue.ProofQueue(provers, ["k0", "k1"], ["v0", "v1"])
    xs = list(range(20, 40))
    res = q.map([_w(x) for x in xs], [(x + 1, x + 2) for x in xs])
    assert [int(r["proof"]["pi_a"][0]) for r in res] == xs                     # input order kept
    assert all(int(r["proof"]["pi_a"][1]) == x + 1 for r, x in zip(res, xs))   # (r, s) travel with their witness
    assert all(r["solidityProof"]["inputs"] == [str(x), "7"] for r, x in zip(res, xs))
    assert provers[0].calls + provers[1].calls == 20 and provers[0].calls > provers[1].calls > 0   # work stealing
    # a failing prove and a proof the verifier rejects are delivered to their submitter only
    f_bad, f_inv, f_ok = q.submit(_w(999), 1, 2), q.submit(_w(13), 1, 2), q.submit(_w(5), 1, 2)
    with pytest.raises(ValueError):
        f_bad.result()
    with pytest.raises(RuntimeError, match="Invalid proof generated"):
        f_inv.result()
    assert int(f_ok.result()["proof"]["pi_a"][0]) == 5
    q.close()
    with pytest.raises(ValueError):
        proof_queue.ProofQueue(provers, ["k0"])


@pytest.mark.gpu
def test_queue_real_proofs_two_contexts():
    import torch
    from oracle import binfmt as bf
    from oracle import groth16 as g
    from simple_zk_rollups_b200 import prover, synth
    toxic = (1234567891011, 222222222222223, 3333333333333331, 44444444444447, 5555555555555557)
    r1, w = synth.generate(120, 3, seed=41)
    pk, vk, _ = g.setup(r1.to_dicts(), toxic)
    pk_bin = bf.binarify_proving_key(pk)
    devs = [0, 1] if torch.cuda.device_count() > 1 else [0, 0]
    provers = [prover.Groth16Prover(d) for d in devs]
    try:
        keys = [p.load_key(pk_bin) for p in provers]
        vkeys = [p.load_vkey(bf.vk_to_json(vk)) for p in provers]
        q = proof_queue.ProofQueue(provers, keys, vkeys)
        rng = random.Random(3)
        rs = [(rng.randrange(R), rng.randrange(R)) for _ in range(8)]
        res = q.map([bf.binarify_witness(w)] * 8, rs)
        for out, (r, s) in zip(res, rs):
            assert out["proof_bytes"] == g.proof_to_bytes(g.gen_proof(pk, w, r, s)[0])
            assert out["solidityProof"]["inputs"] == [str(x) for x in w[1:4]]
        assert sum(q.proved) == 8
        bad = list(w)
        bad[-1] = (bad[-1] + 1) % R                      # violates the circuit: the self-check must raise
        with pytest.raises(RuntimeError, match="Invalid proof generated"):
            q.submit(bf.binarify_witness(bad), 1, 2).result()
        q.close()
    finally:
        for p in provers:
            p.close()
