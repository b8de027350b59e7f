"""BN254 (alt_bn128) field / curve / pairing arithmetic on Python integers.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it.  The product path is the CUDA library
(`simple_zk_rollups_b200/csrc`, C-ABI in `include/zkr.h`).

PARITY UNPINNED: the reference's prover arithmetic lives in
un-vendored npm packages (websnark@0.0.5, snarkjs@0.1.20 -- pins at
/root/reference/operator/yarn.lock:5674,6750) and the reference commits no golden
proofs.  This file restates the published math; it is pinned against every
constant the reference does hold for this path (tests/golden/reference_constants.json):
  * field moduli        operator/src/utils/binarify.ts:80,87; contracts/contracts/TxVerifier.sol:50,259
  * G1 / G2 generators  contracts/contracts/TxVerifier.sol:24-35
  * the two committed verifying keys (78 G1 + 6 G2 points) TxVerifier.sol:177-255,
    WithdrawVerifier.sol:177-185 -- on-curve, in-subgroup, pairing-bilinear fixtures
"""

# ---------------------------------------------------------------- constants
# base field modulus q  (binarify.ts:80, TxVerifier.sol:50)
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
# scalar field modulus r (binarify.ts:87, TxVerifier.sol:259)
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
# BN parameter x:  q = 36x^4+36x^3+24x^2+6x+1,  r = 36x^4+36x^3+18x^2+6x+1
BN_X = 4965661367192848881
ATE_LOOP = 6 * BN_X + 2

MONT_BITS = 256                      # binarify.ts:82,89 -- R_mont = 2^256
MONT_R = 1 << MONT_BITS

G1_GEN = (1, 2)                      # TxVerifier.sol:24-26
# G2 generator; TxVerifier.sol:30-35 lists each Fq2 coordinate imaginary part first
# ("Encoding of field elements is: X[0] * z + X[1]", TxVerifier.sol:18).
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)
B1 = 3                               # G1: y^2 = x^3 + 3


def fq_inv(a):
    return pow(a, -1, Q)


def fr_inv(a):
    return pow(a, -1, R)


# ---------------------------------------------------------------- Fq2 = Fq[u]/(u^2+1)
def f2_add(a, b):
    return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)


def f2_sub(a, b):
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def f2_neg(a):
    return ((-a[0]) % Q, (-a[1]) % Q)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_sqr(a):
    return ((a[0] + a[1]) * (a[0] - a[1]) % Q, 2 * a[0] * a[1] % Q)


def f2_muls(a, k):
    return (a[0] * k % Q, a[1] * k % Q)


def f2_inv(a):
    n = fq_inv((a[0] * a[0] + a[1] * a[1]) % Q)
    return (a[0] * n % Q, (-a[1]) * n % Q)


def f2_conj(a):
    return (a[0], (-a[1]) % Q)


F2_ZERO = (0, 0)
F2_ONE = (1, 0)
XI = (9, 1)                          # sextic non-residue 9+u
B2 = f2_mul((3, 0), f2_inv(XI))      # twist: y^2 = x^3 + 3/(9+u)


# ---------------------------------------------------------------- generic short-Weierstrass (a=0), Jacobian
class _Curve:
    """Jacobian arithmetic parametrised by the coordinate field.  None = infinity."""

    def __init__(self, add, sub, mul, sqr, neg, inv, zero, one, b):
        self.fadd, self.fsub, self.fmul, self.fsqr = add, sub, mul, sqr
        self.fneg, self.finv, self.zero, self.one, self.b = neg, inv, zero, one, b

    def is_on_curve(self, p):
        if p is None:
            return True
        x, y = p
        return self.fsqr(y) == self.fadd(self.fmul(self.fsqr(x), x), self.b)

    def to_jac(self, p):
        return None if p is None else (p[0], p[1], self.one)

    def to_affine(self, p):
        if p is None:
            return None
        x, y, z = p
        if z == self.zero:
            return None
        zi = self.finv(z)
        zi2 = self.fsqr(zi)
        return (self.fmul(x, zi2), self.fmul(y, self.fmul(zi2, zi)))

    def jdbl(self, p):
        if p is None:
            return None
        x, y, z = p
        if y == self.zero:
            return None
        A = self.fsqr(x)
        B = self.fsqr(y)
        C = self.fsqr(B)
        t = self.fsub(self.fsub(self.fsqr(self.fadd(x, B)), A), C)
        D = self.fadd(t, t)
        E = self.fadd(self.fadd(A, A), A)
        F = self.fsqr(E)
        x3 = self.fsub(F, self.fadd(D, D))
        c8 = self.fadd(C, C)
        c8 = self.fadd(c8, c8)
        c8 = self.fadd(c8, c8)
        y3 = self.fsub(self.fmul(E, self.fsub(D, x3)), c8)
        yz = self.fmul(y, z)
        return (x3, y3, self.fadd(yz, yz))

    def jadd(self, p, q):
        if p is None:
            return q
        if q is None:
            return p
        x1, y1, z1 = p
        x2, y2, z2 = q
        z1z1 = self.fsqr(z1)
        z2z2 = self.fsqr(z2)
        u1 = self.fmul(x1, z2z2)
        u2 = self.fmul(x2, z1z1)
        s1 = self.fmul(y1, self.fmul(z2, z2z2))
        s2 = self.fmul(y2, self.fmul(z1, z1z1))
        if u1 == u2:
            if s1 == s2:
                return self.jdbl(p)
            return None
        h = self.fsub(u2, u1)
        rr = self.fsub(s2, s1)
        hh = self.fsqr(h)
        hhh = self.fmul(h, hh)
        v = self.fmul(u1, hh)
        x3 = self.fsub(self.fsub(self.fsqr(rr), hhh), self.fadd(v, v))
        y3 = self.fsub(self.fmul(rr, self.fsub(v, x3)), self.fmul(s1, hhh))
        z3 = self.fmul(self.fmul(z1, z2), h)
        return (x3, y3, z3)

    def jneg(self, p):
        return None if p is None else (p[0], self.fneg(p[1]), p[2])

    def jmul(self, p, k):
        """double-and-add, MSB first (the snarkjs GCurve.mulScalar structure)."""
        if p is None or k == 0:
            return None
        if k < 0:
            return self.jmul(self.jneg(p), -k)
        acc = None
        for bit in bin(k)[2:]:
            acc = self.jdbl(acc)
            if bit == "1":
                acc = self.jadd(acc, p)
        return acc

    # affine conveniences ------------------------------------------------
    def add(self, p, q):
        return self.to_affine(self.jadd(self.to_jac(p), self.to_jac(q)))

    def neg(self, p):
        return None if p is None else (p[0], self.fneg(p[1]))

    def mul(self, p, k):
        return self.to_affine(self.jmul(self.to_jac(p), k))

    def batch_to_affine(self, pts):
        """Montgomery-trick batch inversion; pts = list of Jacobian / None."""
        idx = [i for i, p in enumerate(pts) if p is not None and p[2] != self.zero]
        pref = []
        acc = self.one
        for i in idx:
            acc = self.fmul(acc, pts[i][2])
            pref.append(acc)
        out = [None] * len(pts)
        if not idx:
            return out
        inv = self.finv(acc)
        for j in range(len(idx) - 1, -1, -1):
            i = idx[j]
            zi = self.fmul(inv, pref[j - 1]) if j else inv
            inv = self.fmul(inv, pts[i][2])
            zi2 = self.fsqr(zi)
            out[i] = (self.fmul(pts[i][0], zi2), self.fmul(pts[i][1], self.fmul(zi2, zi)))
        return out


G1 = _Curve(lambda a, b: (a + b) % Q, lambda a, b: (a - b) % Q, lambda a, b: a * b % Q,
            lambda a: a * a % Q, lambda a: (-a) % Q, fq_inv, 0, 1, B1)
G2 = _Curve(f2_add, f2_sub, f2_mul, f2_sqr, f2_neg, f2_inv, F2_ZERO, F2_ONE, B2)


class FixedBase:
    """Windowed fixed-base multiplication (setup generates ~5n multiples of one generator)."""

    def __init__(self, curve, gen, wbits=8, nbits=254):
        self.c, self.w = curve, wbits
        self.nwin = (nbits + wbits - 1) // wbits
        self.tab = []
        base = curve.to_jac(gen)
        for _ in range(self.nwin):
            row = [None]
            acc = None
            for _ in range((1 << wbits) - 1):
                acc = curve.jadd(acc, base)
                row.append(acc)
            aff = curve.batch_to_affine(row)
            self.tab.append([curve.to_jac(p) for p in aff])
            for _ in range(wbits):
                base = curve.jdbl(base)

    def jmul(self, k):
        acc = None
        mask = (1 << self.w) - 1
        for w in range(self.nwin):
            d = (k >> (w * self.w)) & mask
            if d:
                acc = self.c.jadd(acc, self.tab[w][d])
        return acc

    def mul_many(self, ks):
        return self.c.batch_to_affine([self.jmul(k % R) for k in ks])


_fb_cache = {}


def fixed_base(which):
    if which not in _fb_cache:
        _fb_cache[which] = FixedBase(G1, G1_GEN) if which == 1 else FixedBase(G2, G2_GEN)
    return _fb_cache[which]


# ---------------------------------------------------------------- Fq12 = Fq2[w]/(w^6 - xi)
def f12_one():
    return [F2_ONE] + [F2_ZERO] * 5


def f12_mul(a, b):
    t = [[0, 0] for _ in range(11)]
    for i in range(6):
        a0, a1 = a[i]
        if a0 == 0 and a1 == 0:
            continue
        for j in range(6):
            b0, b1 = b[j]
            tt = t[i + j]
            tt[0] += a0 * b0 - a1 * b1
            tt[1] += a0 * b1 + a1 * b0
    out = []
    for k in range(6):
        c0, c1 = t[k]
        if k < 5:
            h0, h1 = t[k + 6]
            c0 += 9 * h0 - h1          # (h0 + h1 u)(9 + u)
            c1 += 9 * h1 + h0
        out.append((c0 % Q, c1 % Q))
    return out


def f12_pow(a, e):
    res = f12_one()
    for bit in bin(e)[2:]:
        res = f12_mul(res, res)
        if bit == "1":
            res = f12_mul(res, a)
    return res


# Frobenius constants on the twist
_G12 = None


def _frob_consts():
    global _G12
    if _G12 is None:
        def f2_pow(a, e):
            r = F2_ONE
            for bit in bin(e)[2:]:
                r = f2_sqr(r)
                if bit == "1":
                    r = f2_mul(r, a)
            return r
        _G12 = (f2_pow(XI, (Q - 1) // 3), f2_pow(XI, (Q - 1) // 2),
                f2_pow(XI, (Q * Q - 1) // 3), f2_pow(XI, (Q * Q - 1) // 2))
    return _G12


def _line(t, q2, p):
    """Line through twist points t, q2 (affine, Fq2) evaluated at P in G1; returns
    (f12 line value, t+q2).  Untwist psi(x,y) = (x w^2, y w^3):
    l(P) = yP - lambda*xP*w + (lambda*x1 - y1)*w^3."""
    x1, y1 = t
    x2, y2 = q2
    if x1 == x2 and y1 == y2:
        lam = f2_mul(f2_muls(f2_sqr(x1), 3), f2_inv(f2_add(y1, y1)))
    elif x1 == x2:
        # vertical line: xP - x1 w^2
        return [(p[0], 0), F2_ZERO, f2_neg(x1), F2_ZERO, F2_ZERO, F2_ZERO], None
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_sqr(lam), x1), x2)
    y3 = f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1)
    ell = [(p[1], 0), f2_neg(f2_muls(lam, p[0])), F2_ZERO,
           f2_sub(f2_mul(lam, x1), y1), F2_ZERO, F2_ZERO]
    return ell, (x3, y3)


def miller_loop(p, q2):
    """Optimal-ate Miller loop f_{6x+2,Q}(P) * two Frobenius lines.  p in G1 affine, q2 in G2 affine."""
    if p is None or q2 is None:
        return f12_one()
    g12, g13, g22, g23 = _frob_consts()
    f = f12_one()
    t = q2
    for bit in bin(ATE_LOOP)[3:]:
        ell, t = _line(t, t, p)
        f = f12_mul(f12_mul(f, f), ell)
        if bit == "1":
            ell, t = _line(t, q2, p)
            f = f12_mul(f, ell)
    q1 = (f2_mul(f2_conj(q2[0]), g12), f2_mul(f2_conj(q2[1]), g13))
    nq2 = (f2_mul(q2[0], g22), f2_neg(f2_mul(q2[1], g23)))
    ell, t = _line(t, q1, p)
    f = f12_mul(f, ell)
    ell, t = _line(t, nq2, p)
    f = f12_mul(f, ell)
    return f


FINAL_EXP = (Q ** 12 - 1) // R


def final_exp(f):
    return f12_pow(f, FINAL_EXP)


def pairing(p, q2):
    return final_exp(miller_loop(p, q2))


def pairing_product_is_one(pairs):
    """prod e(P_i, Q_i) == 1 -- the EVM precompile 8 predicate (TxVerifier.sol:91-116)."""
    f = f12_one()
    for p, q2 in pairs:
        f = f12_mul(f, miller_loop(p, q2))
    return final_exp(f) == f12_one()


def g1_in_subgroup(p):
    return G1.is_on_curve(p) and G1.jmul(G1.to_jac(p), R) is None


def g2_in_subgroup(q2):
    return G2.is_on_curve(q2) and G2.jmul(G2.to_jac(q2), R) is None
