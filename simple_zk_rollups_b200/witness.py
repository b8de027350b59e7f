"""Host side of the GPU witness solver (libzkr zkr_wprog_* / zkr_witness_solve).

Stands where the reference compiles the circuit and runs circom's generated calculator on every proof,
    const circuit = new Circuit(circuitDef); const witness = circuit.calculateWitness(circuitInputs)
(/root/reference/operator/src/snarks/common.ts:12-17), for constraint systems that can be solved forward: every
constraint introduces at most one new signal, in C only (MiMC / Feistel rounds of prover/circuits/hasher.circom:8,
products, linear combinations).  The circuit is analysed once (WitnessSolver), every proof then costs one small upload
of the GIVEN signals (circuit inputs and hints such as constrained bits) and a few hundred tiny launches; the witness
stays resident in HBM and goes straight into zkr_prove_dev.  No oracle, no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


class WitnessSolver:
    """One circuit's witness program on one prover's GPU.

    r1cs: simple_zk_rollups_b200.synth.R1CS (or any object with nVars / nPublic / nConstraints / pool / csc()); the
    circuit's own constraints are used, without the input-consistency rows of a snarkjs setup."""

    def __init__(self, prover, r1cs):
        self.p, self.L = prover, prover.L
        n = r1cs.nVars
        csc = r1cs.csc(with_inputs=False)
        desc = _lib.R1csCsc()
        desc.n_vars, desc.n_public, desc.n_constraints = n, r1cs.nPublic, r1cs.nConstraints
        desc.domain_size, desc.n_pool = r1cs.domain()[1], len(r1cs.pool)
        keep = []
        for k in "abc":
            ptr, row, cid = (np.ascontiguousarray(x, dtype=np.uint32) for x in csc[k.upper()])
            keep += [ptr, row, cid]
            setattr(desc, "ptr_" + k, ptr.ctypes.data)
            setattr(desc, "row_" + k, row.ctypes.data)
            setattr(desc, "cid_" + k, cid.ctypes.data)
        pool = np.frombuffer(b"".join(int(c).to_bytes(32, "little") for c in r1cs.pool), dtype=np.uint8).copy()
        desc.pool = pool.ctypes.data
        self.h = C.c_void_p()
        _lib.check(self.L.zkr_wprog_build(prover.ctx, C.byref(desc), C.byref(self.h)))
        a, b, c, d = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        _lib.check(self.L.zkr_wprog_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        self.n_vars, self.n_given, self.n_solved, self.n_levels = a.value, b.value, c.value, d.value
        g = np.zeros(self.n_given, dtype=np.uint32)
        _lib.check(self.L.zkr_wprog_given(self.h, _lib.buf_ptr(g)))
        self.given_signals = g                      # ascending signal indices the caller must supply
        self.d_witness = C.c_void_p()
        _lib.check(self.L.zkr_dev_malloc(prover.ctx, 32 * self.n_vars, C.byref(self.d_witness)))

    def close(self):
        if self.h:
            self.L.zkr_dev_free(self.p.ctx, self.d_witness)
            self.L.zkr_wprog_free(self.h)
            self.h = C.c_void_p()

    def given_from_witness(self, witness):
        """The GIVEN values picked out of a full witness (list of ints): what a host-side calculator has to produce."""
        return [int(witness[s]) for s in self.given_signals.tolist()]

    def solve_dev(self, given_values):
        """given_values: ints in the order of given_signals (signal 0 must be 1).  -> device pointer (int) of the
        complete witness, n_vars x 32 B standard form, owned by this solver and overwritten by the next call."""
        if len(given_values) != self.n_given:
            raise ValueError("need %d given values, got %d" % (self.n_given, len(given_values)))
        buf = np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in given_values), dtype=np.uint8)
        _lib.check(self.L.zkr_witness_solve(self.p.ctx, self.h, _lib.buf_ptr(buf), self.d_witness))
        return self.d_witness.value

    def solve(self, given_values):
        """-> the complete witness as a list of ints (device solve + download: for tests and for hosts that want it)."""
        self.solve_dev(given_values)
        out = np.empty(32 * self.n_vars, dtype=np.uint8)
        _lib.check(self.L.zkr_dev_download(self.p.ctx, _lib.buf_ptr(out), self.d_witness, out.size))
        b = out.tobytes()
        return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def createGpuProofGenerator(prover, key, solver, n_public, vkey=None):
    """createProofGenerator (common.ts:10-53) with the witness made on the GPU: `given_values` -> witness resident in
    HBM -> zkr_prove_dev -> 256 bytes; only the public signals and the proof cross the bus.  Returns
    generate(given_values, r=None, s=None) -> {"proof": ..., "publicSignals": [...]}; an invalid proof raises
    "Invalid proof generated" (common.ts:36-38) when a verifying-key handle is passed."""
    from .binarify import proof_from_bytes
    from .prover import _random_scalar
    L = prover.L
    d_proof = C.c_void_p()
    _lib.check(L.zkr_dev_malloc(prover.ctx, _lib.PROOF_BYTES, C.byref(d_proof)))

    def generate(given_values, r=None, s=None):
        r = _random_scalar() if r is None else r
        s = _random_scalar() if s is None else s
        d_w = solver.solve_dev(given_values)
        rb = np.frombuffer(int(r).to_bytes(32, "little"), dtype=np.uint8)
        sb = np.frombuffer(int(s).to_bytes(32, "little"), dtype=np.uint8)
        _lib.check(L.zkr_prove_dev(prover.ctx, key, C.c_void_p(d_w), solver.n_vars, _lib.buf_ptr(rb), _lib.buf_ptr(sb), d_proof))
        _lib.check(L.zkr_prove_check(prover.ctx, key))
        out = np.zeros(_lib.PROOF_BYTES, dtype=np.uint8)
        _lib.check(L.zkr_dev_download(prover.ctx, _lib.buf_ptr(out), d_proof, out.size))
        pub = np.zeros(32 * n_public, dtype=np.uint8)
        if n_public:
            _lib.check(L.zkr_dev_download(prover.ctx, _lib.buf_ptr(pub), C.c_void_p(d_w + 32), pub.size))
        pb = pub.tobytes()
        publicSignals = [int.from_bytes(pb[i:i + 32], "little") for i in range(0, len(pb), 32)]
        if vkey is not None and not prover.verify(vkey, out.tobytes(), publicSignals):
            raise RuntimeError("Invalid proof generated")
        return {"proof": proof_from_bytes(out.tobytes()), "proof_bytes": out.tobytes(),
                "publicSignals": [str(x) for x in publicSignals]}

    return generate
