"""Synthetic rollup-shaped R1CS + witness generator (workload generator, host side).

The real circuits (/root/reference/prover/circuits/*.circom, main = BatchProcessTx(2, 6) at
tx.circom:4) cannot be compiled here (no circom / node), so benchmarks and tests use a
synthetic R1CS with the same shape:
  * ~80 % of the constraints of the rollup circuit are MiMC-Feistel rounds
    (hasher.circom:8, MiMCSponge(length, 220, 1)): per round  t = k + xL + c_i,
    t2 = t*t, t4 = t2*t2, xL' = xR + t4*t   -> 3 constraints, 3 new signals,
    A rows with 3/1/1 non-zeros, B rows with 3/1/3, C rows with 1/1/2;
  * ~3 % are Num2Bits-style bit constraints b*(b-1) = 0 (eddsa.circom:29,49,
    processtx.circom:52,60,69) whose witness values are {0,1} -- the MSM digit skew;
  * every constraint introduces exactly one new signal, so nVars = 1 + nPublic + nFree + nConstraints;
  * all inputs are public (batchprocesstx.circom:12-34): nPublic = 1 + B*(18 + 3d).
The witness is obtained by solving the system forward, so A.w * B.w = C.w holds row by row and
proofs made from it verify.  Everything is seeded; nothing here touches the GPU or the oracle.
"""
import random

import numpy as np

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
N_ROUND_CONSTANTS = 220


def rollup_shape(batch, depth):
    """Hand-derived constraint count of BatchProcessTx(batch, depth)."""
    return batch * (21832 + 5303 * depth), 1 + batch * (18 + 3 * depth)


SHAPES = {
    # name: (nConstraints, nPublic)
    "withdraw": (2763, 3),                       # withdraw.circom:25, WithdrawVerifier.sol:181
    "tx": rollup_shape(2, 6),                    # tx.circom:4 -> (107300, 73), m = 2^17
    "tx_2p20": rollup_shape(16, 6),              # (858400, 577), m = 2^20
    "tx_2p22": rollup_shape(64, 6),              # (3433600, 2305), m = 2^22
}


class R1CS:
    """CSC-by-signal sparse A, B, C with coefficients drawn from a small pool."""

    def __init__(self, n_vars, n_public, n_constraints, pool, mats):
        self.nVars, self.nPublic, self.nConstraints = n_vars, n_public, n_constraints
        self.pool = pool                          # list[int] in [0, r)
        self.mats = mats                          # {"A": (sig, row, cid) uint32 arrays sorted by (sig,row)}

    def domain(self):
        need = self.nConstraints + self.nPublic + 1
        bits = max((need - 1).bit_length(), 1)
        return bits, 1 << bits

    def with_input_rows(self):
        """The matrices as snarkjs setup leaves them in the proving key: polsA gains the
        input-consistency rows polsA[i][nConstraints+i] = 1 for i <= nPublic."""
        sig, row, cid = self.mats["A"]
        l, nc = self.nPublic, self.nConstraints
        sig2 = np.concatenate([sig, np.arange(l + 1, dtype=np.uint32)])
        row2 = np.concatenate([row, np.arange(nc, nc + l + 1, dtype=np.uint32)])
        cid2 = np.concatenate([cid, np.zeros(l + 1, dtype=np.uint32)])       # pool[0] == 1
        o = np.lexsort((row2, sig2))
        return {"A": (sig2[o], row2[o], cid2[o]), "B": self.mats["B"], "C": self.mats["C"]}

    def csc(self, with_inputs=True):
        """{"A": (ptr[n+1] uint32, row uint32, cid uint32)}"""
        mats = self.with_input_rows() if with_inputs else self.mats
        out = {}
        for k, (sig, row, cid) in mats.items():
            cnt = np.bincount(sig, minlength=self.nVars).astype(np.uint64)
            ptr = np.zeros(self.nVars + 1, dtype=np.uint64)
            np.cumsum(cnt, out=ptr[1:])
            out[k] = (ptr.astype(np.uint32), row, cid)
        return out

    def to_dicts(self):
        """Per-signal {row: coeff} lists (snarkjs JSON shape, ints) WITHOUT the input rows --
        the form oracle.groth16.setup takes."""
        out = {"nVars": self.nVars, "nPublic": self.nPublic, "nConstraints": self.nConstraints}
        for k, (sig, row, cid) in self.mats.items():
            cols = [dict() for _ in range(self.nVars)]
            for s, r_, c in zip(sig.tolist(), row.tolist(), cid.tolist()):
                cols[s][r_] = self.pool[c]
            out[k] = cols
        return out

    def nnz(self):
        return {k: int(v[0].size) for k, v in self.mats.items()}


def pool_bytes(pool):
    """pool as (len, 32) uint8, little-endian standard form."""
    return np.frombuffer(b"".join(int(c).to_bytes(32, "little") for c in pool), dtype=np.uint8).reshape(-1, 32)


def witness_bytes(witness):
    """binarifyWitness layout (/root/reference/operator/src/utils/binarify.ts:10-48):
    n x 32 bytes, little-endian 8 x u32, standard (non-Montgomery) form."""
    return b"".join(int(x).to_bytes(32, "little") for x in witness)


def generate(n_constraints, n_public, seed=0, n_free=3, bit_every=33, witness_seed=None):
    """-> (R1CS, witness list[int]).  Deterministic in (n_constraints, n_public, seed, witness_seed).
    witness_seed = None: one generator draws structure and input values (the round-1 behaviour every committed fixture
    was made with).  witness_seed = k: the input and bit values come from a second generator, so the same
    (n_constraints, n_public, seed) gives the SAME circuit for every k and a different satisfying witness per k
    (distinct proofs under one proving key: BASELINE.json configs[4])."""
    rng = random.Random((0x7A6B726F6C6C7570 ^ seed) & 0xFFFFFFFFFFFFFFFF)
    wrng = rng if witness_seed is None else random.Random((0x7769746E65737300 ^ (seed << 20) ^ witness_seed) & 0xFFFFFFFFFFFFFFFF)
    pool = [1, R - 1] + [rng.randrange(R) for _ in range(N_ROUND_CONSTANTS)]
    ONE, P_ONE, P_NEG = 0, 0, 1
    w = [1] + [wrng.randrange(R) for _ in range(n_public + n_free)]
    n_in = n_public + n_free
    ent = {"A": ([], [], []), "B": ([], [], []), "C": ([], [], [])}

    def put(mat, row, sig, cid):
        e = ent[mat]
        e[0].append(sig)
        e[1].append(row)
        e[2].append(cid)

    start_ctr = 0

    def restart():
        nonlocal start_ctr
        # cycle through the inputs so every public signal is constrained; three distinct signals
        a = 1 + (start_ctr % n_in)
        b = 1 + ((start_ctr + 1) % n_in)
        c = 1 + ((start_ctr + 2) % n_in)
        start_ctr += 3
        return a, b, c

    k, xl, xr = restart()
    rnd = 0
    row = 0
    while row < n_constraints:
        if bit_every and row % bit_every == bit_every - 1:
            b = len(w)
            w.append(wrng.getrandbits(1))
            put("A", row, b, P_ONE)
            put("B", row, ONE, P_NEG)
            put("B", row, b, P_ONE)
            row += 1
            continue
        if n_constraints - row < 3 or (bit_every and (row % bit_every) > bit_every - 4):
            # filler: plain product of two earlier signals
            i = 1 + rng.randrange(len(w) - 1)
            j = 1 + rng.randrange(len(w) - 1)
            v = len(w)
            w.append(w[i] * w[j] % R)
            put("A", row, i, P_ONE)
            put("B", row, j, P_ONE)
            put("C", row, v, P_ONE)
            row += 1
            continue
        ci = 2 + (rnd % N_ROUND_CONSTANTS)
        t = (w[k] + w[xl] + pool[ci]) % R
        t2s, t4s, xns = len(w), len(w) + 1, len(w) + 2
        t2 = t * t % R
        t4 = t2 * t2 % R
        xn = (t4 * t + w[xr]) % R
        w.extend((t2, t4, xn))
        for mat in ("A", "B"):
            put(mat, row, ONE, ci)
            put(mat, row, k, P_ONE)
            put(mat, row, xl, P_ONE)
        put("C", row, t2s, P_ONE)
        put("A", row + 1, t2s, P_ONE)
        put("B", row + 1, t2s, P_ONE)
        put("C", row + 1, t4s, P_ONE)
        put("A", row + 2, t4s, P_ONE)
        put("B", row + 2, ONE, ci)
        put("B", row + 2, k, P_ONE)
        put("B", row + 2, xl, P_ONE)
        put("C", row + 2, xns, P_ONE)
        put("C", row + 2, xr, P_NEG)
        row += 3
        rnd += 1
        xr, xl = xl, xns
        if rnd % N_ROUND_CONSTANTS == 0:
            k, xl, xr = restart()
    mats = {}
    for name, (sig, rw, cid) in ent.items():
        sig = np.asarray(sig, dtype=np.uint32)
        rw = np.asarray(rw, dtype=np.uint32)
        cid = np.asarray(cid, dtype=np.uint32)
        o = np.lexsort((rw, sig))
        mats[name] = (sig[o], rw[o], cid[o])
    r1cs = R1CS(len(w), n_public, n_constraints, pool, mats)
    assert r1cs.nVars == 1 + n_public + n_free + n_constraints
    return r1cs, w


def check_witness(r1cs, w):
    """A.w * B.w == C.w on every row (host check, Python ints)."""
    acc = {}
    for name, (sig, row, cid) in r1cs.mats.items():
        v = [0] * r1cs.nConstraints
        for s, r_, c in zip(sig.tolist(), row.tolist(), cid.tolist()):
            v[r_] = (v[r_] + w[s] * r1cs.pool[c]) % R
        acc[name] = v
    return all(a * b % R == c for a, b, c in zip(acc["A"], acc["B"], acc["C"]))
