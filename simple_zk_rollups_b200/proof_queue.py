"""Multi-GPU proof queue: the consumer side of the operator batch loop the reference stubs out.

In the reference, `POST /send` validates a transaction and pushes it on a redis list
(operator/src/routes/send.ts:142-147); the batch loop that would drain `batchSize` transactions
(zk-rollups.config.js:31-34), build the BatchProcessTx inputs, prove and call `rollUp` is absent -- the only
place the prove is driven is the test script operator/__tests__/operatorLogic.test.ts:105-231.  What that loop
needs from this repository is exactly this: witnesses in, (proof, solidityProof) out, as many in flight as there
are GPUs, each self-checked like common.ts:30-38.  Queueing, redis, the DB and the contract call stay the
operator's (only the prove queue is in scope).

One worker thread per GPU context; each owns a resident proving key (and verifying key) on its device and calls
zkr_prove / zkr_verify -- no CPU fallback, no oracle.  Results are delivered through concurrent.futures.Future in
submission order per caller; `map` keeps input order.
"""
import concurrent.futures
import queue
import threading

from . import _lib
from .binarify import R as SNARK_FIELD_SIZE
from .binarify import proof_from_bytes


class ProofQueue:
    """provers: list of prover.Groth16Prover (one per GPU); keys[i] / vkeys[i]: handles of the SAME circuit's
    proving / verifying key loaded on provers[i] (vkeys optional: no self-check without them)."""

    def __init__(self, provers, keys, vkeys=None, n_public=None):
        if len(provers) != len(keys) or (vkeys is not None and len(vkeys) != len(provers)):
            raise ValueError("one key (and verifying key) per prover")
        self.provers, self.keys, self.vkeys = provers, keys, vkeys
        self.n_public = provers[0].key_info(keys[0])["nPublic"] if n_public is None else n_public
        self._q = queue.Queue()
        self._threads = [threading.Thread(target=self._work, args=(i,), daemon=True) for i in range(len(provers))]
        self.proved = [0] * len(provers)
        for t in self._threads:
            t.start()

    def submit(self, witness_bin, r, s):
        """-> Future of {"proof": {pi_a, pi_b, pi_c, protocol}, "solidityProof": {a, b, c, inputs}, "gpu": i}.
        r, s: blinding scalars (ints < r; the caller draws them, websnark draws them internally)."""
        f = concurrent.futures.Future()
        self._q.put((f, witness_bin, int(r), int(s)))
        return f

    def map(self, witness_bins, rs):
        futs = [self.submit(w, r, s) for w, (r, s) in zip(witness_bins, rs)]
        return [f.result() for f in futs]

    def close(self):
        for _ in self._threads:
            self._q.put(None)
        for t in self._threads:
            t.join()

    def _work(self, i):
        p, key = self.provers[i], self.keys[i]
        vkey = self.vkeys[i] if self.vkeys is not None else None
        while True:
            job = self._q.get()
            if job is None:
                return
            f, wbin, r, s = job
            if not f.set_running_or_notify_cancel():
                continue
            try:
                buf, stats = p.prove(key, wbin, r, s)
                wb = bytes(memoryview(wbin)[32:32 * (self.n_public + 1)])       # public signals = witness[1 .. nPublic]
                pub = [int.from_bytes(wb[32 * k:32 * k + 32], "little") for k in range(self.n_public)]
                if vkey is not None and not p.verify(vkey, buf, pub):
                    raise RuntimeError("Invalid proof generated")          # common.ts:36-38
                proof = proof_from_bytes(buf)
                self.proved[i] += 1
                f.set_result({
                    "proof": proof,
                    "solidityProof": {"a": proof["pi_a"][:2], "b": [list(reversed(x)) for x in proof["pi_b"]][:2],
                                      "c": proof["pi_c"][:2], "inputs": [str(x % SNARK_FIELD_SIZE) for x in pub]},
                    "proof_bytes": buf, "gpu": p.device, "prove_ms": stats["total_ms"]})
            except (Exception, _lib.ZkrError) as e:      # noqa: BLE001 -- delivered to the submitter
                f.set_exception(e)
