"""Groth16 over BN254 on Python integers: NTT, H, setup, prove, verify.

TEST INFRASTRUCTURE ONLY (see oracle/bn254.py header).  PARITY UNPINNED: restates
  * snarkjs@0.1.20 src/prover_groth.js / polfield.js / setup_groth.js / verifier_groth.js
  * websnark@0.0.5 src/groth16.js (the prover the reference actually calls:
    /root/reference/operator/src/snarks/common.ts:29)
from their published algorithms; neither package is vendored.
The acceptance predicate follows /root/reference/contracts/contracts/TxVerifier.sol:258-276.

Key / proof containers mirror the snarkjs JSON schema consumed by
/root/reference/operator/src/utils/binarify.ts:129-202 but hold ints / tuples:
  pk = {nVars, nPublic, domainBits, domainSize, polsA/B/C: [ {row: coeff} ], A/B1/C/hExps: [(x,y)|None],
        B2: [((x0,x1),(y0,y1))|None], vk_alfa_1, vk_beta_1, vk_delta_1, vk_beta_2, vk_delta_2}
  vk = {nPublic, IC, vk_alfa_1, vk_beta_2, vk_gamma_2, vk_delta_2}
"""
from . import bn254 as bn
from .bn254 import R, G1, G2

# ---------------------------------------------------------------- roots of unity
TWO_ADICITY = 28
assert (R - 1) % (1 << TWO_ADICITY) == 0 and (R - 1) % (1 << (TWO_ADICITY + 1)) != 0


def root_of_unity(bits):
    """omega_k = 5^((r-1)/2^k): snarkjs PolField uses the smallest non-residue (5) as generator."""
    assert 0 <= bits <= TWO_ADICITY
    return pow(5, (R - 1) >> bits, R)


def _fft_rec(a, w):
    """Recursive radix-2 (the snarkjs polfield.js __fft structure): natural in, natural out."""
    n = len(a)
    if n == 1:
        return a
    w2 = w * w % R
    ev = _fft_rec(a[0::2], w2)
    od = _fft_rec(a[1::2], w2)
    out = [0] * n
    t = 1
    h = n // 2
    for i in range(h):
        x = od[i] * t % R
        out[i] = (ev[i] + x) % R
        out[i + h] = (ev[i] - x) % R
        t = t * w % R
    return out


def ntt(a, inverse=False):
    n = len(a)
    bits = n.bit_length() - 1
    assert 1 << bits == n
    w = root_of_unity(bits)
    if inverse:
        w = pow(w, -1, R)
    out = _fft_rec(list(a), w)
    if inverse:
        ninv = pow(n, -1, R)
        out = [x * ninv % R for x in out]
    return out


def coset_ntt(a, shift):
    """evaluate poly a on shift*<omega>."""
    s = 1
    b = []
    for x in a:
        b.append(x * s % R)
        s = s * shift % R
    return ntt(b)


def coset_intt(e, shift):
    a = ntt(e, inverse=True)
    si = pow(shift, -1, R)
    s = 1
    out = []
    for x in a:
        out.append(x * s % R)
        s = s * si % R
    return out


def bit_reverse(i, bits):
    return int(bin(i)[2:].zfill(bits)[::-1], 2) if bits else 0


# ---------------------------------------------------------------- H = (A*B - C)/Z
def eval_lc(pols, witness, m):
    """A_T[c] = sum_s w_s * pols[s][c]  (websnark pol_constructLC / snarkjs calculateH loop)."""
    out = [0] * m
    for s, col in enumerate(pols):
        ws = witness[s]
        if ws == 0:
            continue
        for c, coef in col.items():
            out[c] = (out[c] + ws * coef) % R
    return out


def calc_h_snarkjs(pk, witness):
    """snarkjs prover_groth.js calculateH: iFFT A,B,C -> poly mul -> sub -> slice(m).  Uses polsC."""
    m = pk["domainSize"]
    a = ntt(eval_lc(pk["polsA"], witness, m), inverse=True)
    b = ntt(eval_lc(pk["polsB"], witness, m), inverse=True)
    c = ntt(eval_lc(pk["polsC"], witness, m), inverse=True)
    # product via size-2m NTT
    fa = ntt(a + [0] * m)
    fb = ntt(b + [0] * m)
    ab = ntt([x * y % R for x, y in zip(fa, fb)], inverse=True)
    p = [(ab[i] - (c[i] if i < m else 0)) % R for i in range(2 * m)]
    return p[m:], p[:m]          # (h, low part; P = H (x^m - 1) so low == -h iff the witness is valid)


def calc_h_websnark(pk, witness):
    """websnark groth16.js calcH: A,B on the 2m domain (even slots given, odd slots via
    iNTT_m + shifted NTT_m), pointwise multiply, iNTT_2m, upper half.  polsC is not used
    (binarify.ts never serialises it)."""
    m = pk["domainSize"]
    bits = m.bit_length() - 1
    g = root_of_unity(bits + 1)
    at = eval_lc(pk["polsA"], witness, m)
    bt = eval_lc(pk["polsB"], witness, m)
    ao = coset_ntt(ntt(at, inverse=True), g)
    bo = coset_ntt(ntt(bt, inverse=True), g)
    ab2 = [0] * (2 * m)
    for i in range(m):
        ab2[2 * i] = at[i] * bt[i] % R
        ab2[2 * i + 1] = ao[i] * bo[i] % R
    coef = ntt(ab2, inverse=True)
    return coef[m:]


def calc_h_coset(pk, witness):
    """method (iii): evaluate A,B,C on the coset g<omega>, divide by Z = g^m - 1 = -2."""
    m = pk["domainSize"]
    bits = m.bit_length() - 1
    g = root_of_unity(bits + 1)
    at = eval_lc(pk["polsA"], witness, m)
    bt = eval_lc(pk["polsB"], witness, m)
    ct = [x * y % R for x, y in zip(at, bt)]
    ac = coset_ntt(ntt(at, inverse=True), g)
    bc = coset_ntt(ntt(bt, inverse=True), g)
    cc = coset_ntt(ntt(ct, inverse=True), g)
    zi = pow(R - 2, -1, R)
    return coset_intt([(x * y - z) * zi % R for x, y, z in zip(ac, bc, cc)], g)


def calc_h_lu(pk, witness):
    """method (iv), the one the CUDA path uses: A*B = L + x^m U;
    iNTT(A_T.B_T) = L+U, coset-iNTT(A.B on coset) = L-U, h = U = ((L+U)-(L-U))/2."""
    m = pk["domainSize"]
    bits = m.bit_length() - 1
    g = root_of_unity(bits + 1)
    at = eval_lc(pk["polsA"], witness, m)
    bt = eval_lc(pk["polsB"], witness, m)
    lpu = ntt([x * y % R for x, y in zip(at, bt)], inverse=True)
    ac = coset_ntt(ntt(at, inverse=True), g)
    bc = coset_ntt(ntt(bt, inverse=True), g)
    lmu = coset_intt([x * y % R for x, y in zip(ac, bc)], g)
    i2 = pow(2, -1, R)
    return [(x - y) * i2 % R for x, y in zip(lpu, lmu)]


# ---------------------------------------------------------------- setup (snarkjs setup_groth.js restated)
def lagrange_at(m, t):
    """L_c(t) for c < m on the domain <omega_m>:  L_c(t) = omega^c (t^m - 1) / (m (t - omega^c))."""
    bits = m.bit_length() - 1
    w = root_of_unity(bits)
    zt = (pow(t, m, R) - 1) % R
    k = zt * pow(m, -1, R) % R
    out = []
    wc = 1
    for _ in range(m):
        out.append(k * wc % R * pow((t - wc) % R, -1, R) % R)
        wc = wc * w % R
    return out


def domain_size(n_constraints, n_public):
    need = n_constraints + n_public + 1
    bits = max((need - 1).bit_length(), 1)
    return bits, 1 << bits


def setup(r1cs, toxic):
    """r1cs = {nVars, nPublic, nConstraints, A/B/C: per-signal [ {row: coeff} ]}  ->  (pk, vk, secrets)
    toxic = (tau, alpha, beta, gamma, delta), all nonzero mod r."""
    tau, alpha, beta, gamma, delta = [x % R for x in toxic]
    n, l, nc = r1cs["nVars"], r1cs["nPublic"], r1cs["nConstraints"]
    bits, m = domain_size(nc, l)
    polsA = [dict(d) for d in r1cs["A"]]
    polsB = [dict(d) for d in r1cs["B"]]
    polsC = [dict(d) for d in r1cs["C"]]
    for i in range(l + 1):                      # input-consistency rows
        polsA[i][nc + i] = 1
    lag = lagrange_at(m, tau)

    def ev(col):
        return sum(c * lag[row] for row, c in col.items()) % R

    a_t = [ev(polsA[i]) for i in range(n)]
    b_t = [ev(polsB[i]) for i in range(n)]
    c_t = [ev(polsC[i]) for i in range(n)]
    dinv, ginv = pow(delta, -1, R), pow(gamma, -1, R)
    k_t = [(beta * a_t[i] + alpha * b_t[i] + c_t[i]) % R for i in range(n)]
    zt = (pow(tau, m, R) - 1) % R
    f1, f2 = bn.fixed_base(1), bn.fixed_base(2)
    A = f1.mul_many(a_t)
    B1 = f1.mul_many(b_t)
    B2 = f2.mul_many(b_t)
    Cp = f1.mul_many([k_t[i] * dinv for i in range(l + 1, n)])
    IC = f1.mul_many([k_t[i] * ginv for i in range(l + 1)])
    hs, tp = [], 1
    for _ in range(m):
        hs.append(tp * zt % R * dinv % R)
        tp = tp * tau % R
    hExps = f1.mul_many(hs)
    g1s = f1.mul_many([alpha, beta, delta])
    g2s = f2.mul_many([beta, gamma, delta])
    pk = dict(protocol="groth", nVars=n, nPublic=l, domainBits=bits, domainSize=m,
              polsA=polsA, polsB=polsB, polsC=polsC, A=A, B1=B1, B2=B2,
              C=[None] * (l + 1) + Cp, hExps=hExps,
              vk_alfa_1=g1s[0], vk_beta_1=g1s[1], vk_delta_1=g1s[2],
              vk_beta_2=g2s[0], vk_delta_2=g2s[2])
    vk = dict(protocol="groth", nPublic=l, IC=IC, vk_alfa_1=g1s[0], vk_beta_2=g2s[0],
              vk_gamma_2=g2s[1], vk_delta_2=g2s[2])
    secrets = dict(tau=tau, alpha=alpha, beta=beta, gamma=gamma, delta=delta,
                   a_t=a_t, b_t=b_t, c_t=c_t, zt=zt)
    return pk, vk, secrets


# ---------------------------------------------------------------- prove
def msm_naive(curve, points, scalars):
    """sum k_i P_i by per-point double-and-add (the snarkjs genProof loop structure)."""
    acc = None
    for p, k in zip(points, scalars):
        if p is None or k % R == 0:
            continue
        acc = curve.jadd(acc, curve.jmul(curve.to_jac(p), k % R))
    return acc


def gen_proof(pk, witness, r=0, s=0, h_method=calc_h_lu):
    """SURVEY Appendix B.2.  r = s = 0 is the snarkjs debug mode; any fixed (r, s) gives a unique proof."""
    n, l = pk["nVars"], pk["nPublic"]
    assert len(witness) == n and witness[0] == 1
    w = [x % R for x in witness]
    h = h_method(pk, w)
    if isinstance(h, tuple):
        h = h[0]
    J1, J2 = G1.to_jac, G2.to_jac
    sa = msm_naive(G1, pk["A"], w)
    sb1 = msm_naive(G1, pk["B1"], w)
    sb2 = msm_naive(G2, pk["B2"], w)
    sc = msm_naive(G1, pk["C"][l + 1:], w[l + 1:])
    sh = msm_naive(G1, pk["hExps"], h)
    d1 = J1(pk["vk_delta_1"])
    pi_a = G1.jadd(G1.jadd(J1(pk["vk_alfa_1"]), sa), G1.jmul(d1, r))
    pi_b = G2.jadd(G2.jadd(J2(pk["vk_beta_2"]), sb2), G2.jmul(J2(pk["vk_delta_2"]), s))
    pib1 = G1.jadd(G1.jadd(J1(pk["vk_beta_1"]), sb1), G1.jmul(d1, s))
    pi_c = G1.jadd(sc, sh)
    pi_c = G1.jadd(pi_c, G1.jmul(pi_a, s))
    pi_c = G1.jadd(pi_c, G1.jmul(pib1, r))
    pi_c = G1.jadd(pi_c, G1.jneg(G1.jmul(d1, r * s % R)))
    proof = dict(pi_a=G1.to_affine(pi_a), pi_b=G2.to_affine(pi_b), pi_c=G1.to_affine(pi_c), protocol="groth")
    return proof, w[1:l + 1]


def proof_to_bytes(proof):
    """The 256-byte C-ABI proof layout (include/zkr.h): pi_a x|y, pi_b x.c0|x.c1|y.c0|y.c1, pi_c x|y;
    32-byte little-endian, standard (non-Montgomery) form, affine; infinity = all zero."""
    def e(v):
        return int(v).to_bytes(32, "little")
    a, b, c = proof["pi_a"], proof["pi_b"], proof["pi_c"]
    out = b""
    out += (e(a[0]) + e(a[1])) if a else bytes(64)
    out += (e(b[0][0]) + e(b[0][1]) + e(b[1][0]) + e(b[1][1])) if b else bytes(128)
    out += (e(c[0]) + e(c[1])) if c else bytes(64)
    return out


def proof_from_bytes(buf):
    v = [int.from_bytes(buf[i * 32:(i + 1) * 32], "little") for i in range(8)]
    z = lambda *xs: all(x == 0 for x in xs)
    return dict(pi_a=None if z(v[0], v[1]) else (v[0], v[1]),
                pi_b=None if z(*v[2:6]) else ((v[2], v[3]), (v[4], v[5])),
                pi_c=None if z(v[6], v[7]) else (v[6], v[7]), protocol="groth")


# ---------------------------------------------------------------- verify
def verify(vk, proof, public_signals):
    """TxVerifier.sol:258-276: inputs < r; vk_x = IC0 + sum in_i IC_{i+1};
    e(-A,B) e(alfa1,beta2) e(vk_x,gamma2) e(C,delta2) == 1."""
    if len(public_signals) + 1 != len(vk["IC"]):
        return False
    if any(not (0 <= x < R) for x in public_signals):
        return False
    a, b, c = proof["pi_a"], proof["pi_b"], proof["pi_c"]
    if a is None or b is None or c is None:
        return False
    if not (G1.is_on_curve(a) and G1.is_on_curve(c) and G2.is_on_curve(b)):
        return False
    vkx = G1.to_jac(vk["IC"][0])
    for x, p in zip(public_signals, vk["IC"][1:]):
        vkx = G1.jadd(vkx, G1.jmul(G1.to_jac(p), x))
    vkx = G1.to_affine(vkx)
    return bn.pairing_product_is_one([
        (G1.neg(a), b), (vk["vk_alfa_1"], vk["vk_beta_2"]),
        (vkx, vk["vk_gamma_2"]), (c, vk["vk_delta_2"])])


def exponent_check(pk, secrets, witness, proof, r=0, s=0):
    """Toxic-waste check (SURVEY 8c leg 2): with (tau,alpha,beta,gamma,delta) known the proof's
    discrete logs are computable in Fr with no MSM / NTT code at all:
       a = alpha + sum w_i A_i(tau) + r delta,  b = beta + sum w_i B_i(tau) + s delta,
       c = (sum_{i>l} w_i K_i(tau) + H(tau) Z(tau)) / delta + s a + r b - r s delta,
    where H(tau) Z(tau) = A(tau) B(tau) - C(tau) for a valid witness."""
    n, l = pk["nVars"], pk["nPublic"]
    w = [x % R for x in witness]
    S = secrets
    at = sum(w[i] * S["a_t"][i] for i in range(n)) % R
    bt = sum(w[i] * S["b_t"][i] for i in range(n)) % R
    ct = sum(w[i] * S["c_t"][i] for i in range(n)) % R
    a = (S["alpha"] + at + r * S["delta"]) % R
    b = (S["beta"] + bt + s * S["delta"]) % R
    kpriv = sum(w[i] * (S["beta"] * S["a_t"][i] + S["alpha"] * S["b_t"][i] + S["c_t"][i])
                for i in range(l + 1, n)) % R
    hz = (at * bt - ct) % R
    c = ((kpriv + hz) * pow(S["delta"], -1, R) + s * a + r * b - r * s * S["delta"]) % R
    f1, f2 = bn.fixed_base(1), bn.fixed_base(2)
    return (proof["pi_a"] == f1.mul_many([a])[0] and proof["pi_b"] == f2.mul_many([b])[0]
            and proof["pi_c"] == f1.mul_many([c])[0])


def exponent_sums_flat(mats, pool, n_public, witness, toxic, m):
    """The sparse pass of exponent_check_flat: per matrix, tot = sum over entries of w[sig] coeff L_row(tau) and
    priv = the same over the private signals.  Independent of (r, s): compute once, check several proofs."""
    tau = toxic[0] % R
    lag = lagrange_at(m, tau)
    w = [x % R for x in witness]
    tot, priv = {}, {}
    for name, (sig, row, cid) in mats.items():
        t = p = 0
        for sg, rw, c in zip(sig.tolist(), row.tolist(), cid.tolist()):
            v = w[sg] * pool[c] % R * lag[rw]
            t += v
            if sg > n_public:
                p += v
        tot[name], priv[name] = t % R, p % R
    return tot, priv


def exponent_check_flat(mats, pool, n_public, witness, toxic, m, proof, r=0, s=0, sums=None):
    """exponent_check for large synthetic circuits: one pass over the sparse entries, no per-signal
    tables.  mats = {"A": (sig, row, cid)} WITH the input-consistency rows in A; pool = coefficient ints.
    sums = (tot, priv) from exponent_sums_flat, or from its C restatement oracle.cbind.exponent_sums (2^22-sized
    circuits; tests/test_oracle_c.py holds the two equal)."""
    tau, alpha, beta, gamma, delta = [x % R for x in toxic]
    tot, priv = sums if sums is not None else exponent_sums_flat(mats, pool, n_public, witness, toxic, m)
    a = (alpha + tot["A"] + r * delta) % R
    b = (beta + tot["B"] + s * delta) % R
    kpriv = (beta * priv["A"] + alpha * priv["B"] + priv["C"]) % R
    hz = (tot["A"] * tot["B"] - tot["C"]) % R
    c = ((kpriv + hz) * pow(delta, -1, R) + s * a + r * b - r * s * delta) % R
    f1, f2 = bn.fixed_base(1), bn.fixed_base(2)
    return (proof["pi_a"] == f1.mul_many([a])[0] and proof["pi_b"] == f2.mul_many([b])[0]
            and proof["pi_c"] == f1.mul_many([c])[0])
