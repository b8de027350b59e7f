"""Restatement of /root/reference/operator/src/utils/binarify.ts (the websnark wire formats).

TEST INFRASTRUCTURE ONLY (see oracle/bn254.py header).  Follows, line by line:
  binarifyWitness      binarify.ts:10-48    n x 32 B, 8 x u32 little-endian, standard form
  binarifyProvingKey   binarify.ts:50-207   header (:152-161), vk points (:163-167), polsA/polsB
                                            (:169-177), A, B1, B2, C, hExps (:179-202);
                                            coordinates x 2^256 mod q (:78-83), coefficients x 2^256 mod r (:85-90)
plus the inverse parser and the snarkjs-JSON <-> int-container conversion
(stringifybigint semantics used at common.ts:17,31-33,44-49).
"""
import struct

from .bn254 import Q, R, MONT_R


def _le32(v):
    return int(v).to_bytes(32, "little")


def to_mont_q(v):          # binarify.ts:78-83
    return v * MONT_R % Q


def to_mont_r(v):          # binarify.ts:85-90
    return v * MONT_R % R


def binarify_witness(witness):
    return b"".join(_le32(x) for x in witness)


def _pt1(p):
    """writePoint (binarify.ts:92-95): z is dropped; snarkjs affine zero is [0,1,0] -> (0, R mod q)."""
    x, y = (0, 1) if p is None else p
    return _le32(to_mont_q(x)) + _le32(to_mont_q(y))


def _pt2(p):
    """writePoint2 (binarify.ts:97-102); G2 zero is [[0,0],[1,0],[0,0]]."""
    (x0, x1), (y0, y1) = ((0, 0), (1, 0)) if p is None else p
    return b"".join(_le32(to_mont_q(v)) for v in (x0, x1, y0, y1))


def _pol(col):
    """writeTransformedPolynomial (binarify.ts:104-113); Object.keys order of integer-like keys
    is ascending numeric."""
    out = [struct.pack("<I", len(col))]
    for row in sorted(col):
        out.append(struct.pack("<I", row))
        out.append(_le32(to_mont_r(col[row])))
    return b"".join(out)


def binarify_proving_key(pk):
    n, l, m = pk["nVars"], pk["nPublic"], pk["domainSize"]
    body = [_pt1(pk["vk_alfa_1"]), _pt1(pk["vk_beta_1"]), _pt1(pk["vk_delta_1"]),
            _pt2(pk["vk_beta_2"]), _pt2(pk["vk_delta_2"])]
    off = 40 + sum(len(b) for b in body)
    ptrs = []

    def section(chunks):
        nonlocal off
        ptrs.append(off)
        blob = b"".join(chunks)
        off += len(blob)
        body.append(blob)

    section(_pol(pk["polsA"][i]) for i in range(n))
    section(_pol(pk["polsB"][i]) for i in range(n))
    section(_pt1(pk["A"][i]) for i in range(n))
    section(_pt1(pk["B1"][i]) for i in range(n))
    section(_pt2(pk["B2"][i]) for i in range(n))
    section(_pt1(pk["C"][i]) for i in range(l + 1, n))
    section(_pt1(pk["hExps"][i]) for i in range(m))
    head = struct.pack("<10I", n, l, m, *ptrs)
    out = head + b"".join(body)
    # calculateBuffLen (binarify.ts:115-141)
    expect = 40 + 3 * 64 + 2 * 128 + sum(36 * len(pk["polsA"][i]) + 4 + 36 * len(pk["polsB"][i]) + 4
                                         for i in range(n)) + n * 64 * 2 + n * 128 + (n - l - 1) * 64 + m * 64
    assert len(out) == expect
    return out


def parse_proving_key(buf):
    """binary -> int container (polsC absent: binarify never writes it)."""
    n, l, m, pA, pB, pPA, pPB1, pPB2, pPC, pPH = struct.unpack_from("<10I", buf, 0)
    rinv_q = pow(MONT_R, -1, Q)
    rinv_r = pow(MONT_R, -1, R)

    def fq(o):
        return int.from_bytes(buf[o:o + 32], "little") * rinv_q % Q

    def pt1(o):
        x, y = fq(o), fq(o + 32)
        return None if x == 0 else (x, y)            # loader rule: x == 0 <=> infinity

    def pt2(o):
        v = [fq(o + 32 * i) for i in range(4)]
        return None if v[0] == 0 and v[1] == 0 else ((v[0], v[1]), (v[2], v[3]))

    def pols(o):
        out = []
        for _ in range(n):
            (k,) = struct.unpack_from("<I", buf, o)
            o += 4
            col = {}
            for _ in range(k):
                (row,) = struct.unpack_from("<I", buf, o)
                col[row] = int.from_bytes(buf[o + 4:o + 36], "little") * rinv_r % R
                o += 36
            out.append(col)
        return out, o

    polsA, endA = pols(pA)
    polsB, endB = pols(pB)
    assert endA == pB and endB == pPA
    pk = dict(protocol="groth", nVars=n, nPublic=l, domainSize=m, domainBits=m.bit_length() - 1,
              polsA=polsA, polsB=polsB,
              vk_alfa_1=pt1(40), vk_beta_1=pt1(104), vk_delta_1=pt1(168),
              vk_beta_2=pt2(232), vk_delta_2=pt2(360),
              A=[pt1(pPA + 64 * i) for i in range(n)],
              B1=[pt1(pPB1 + 64 * i) for i in range(n)],
              B2=[pt2(pPB2 + 128 * i) for i in range(n)],
              C=[None] * (l + 1) + [pt1(pPC + 64 * i) for i in range(n - l - 1)],
              hExps=[pt1(pPH + 64 * i) for i in range(m)])
    assert pPH + 64 * m == len(buf)
    return pk


# ---------------------------------------------------------------- snarkjs JSON schema (decimal strings)
def _s(v):
    return str(int(v))


def g1_to_json(p):
    return ["0", "1", "0"] if p is None else [_s(p[0]), _s(p[1]), "1"]


def g2_to_json(p):
    if p is None:
        return [["0", "0"], ["1", "0"], ["0", "0"]]
    return [[_s(p[0][0]), _s(p[0][1])], [_s(p[1][0]), _s(p[1][1])], ["1", "0"]]


def g1_from_json(j):
    return None if j is None or int(j[2]) == 0 else (int(j[0]), int(j[1]))


def g2_from_json(j):
    if j is None or (int(j[2][0]) == 0 and int(j[2][1]) == 0):
        return None
    return ((int(j[0][0]), int(j[0][1])), (int(j[1][0]), int(j[1][1])))


def pk_to_json(pk):
    l = pk["nPublic"]
    out = {k: pk[k] for k in ("protocol", "nVars", "nPublic", "domainBits", "domainSize")}
    for k in ("polsA", "polsB", "polsC"):
        if k in pk:
            out[k] = [{str(r_): _s(c) for r_, c in sorted(col.items())} for col in pk[k]]
    for k in ("A", "B1", "hExps"):
        out[k] = [g1_to_json(p) for p in pk[k]]
    out["C"] = [None] * (l + 1) + [g1_to_json(p) for p in pk["C"][l + 1:]]
    out["B2"] = [g2_to_json(p) for p in pk["B2"]]
    for k in ("vk_alfa_1", "vk_beta_1", "vk_delta_1"):
        out[k] = g1_to_json(pk[k])
    for k in ("vk_beta_2", "vk_delta_2"):
        out[k] = g2_to_json(pk[k])
    return out


def vk_to_json(vk):
    return dict(protocol="groth", nPublic=vk["nPublic"], IC=[g1_to_json(p) for p in vk["IC"]],
                vk_alfa_1=g1_to_json(vk["vk_alfa_1"]), vk_beta_2=g2_to_json(vk["vk_beta_2"]),
                vk_gamma_2=g2_to_json(vk["vk_gamma_2"]), vk_delta_2=g2_to_json(vk["vk_delta_2"]))


def vk_from_json(j):
    return dict(protocol="groth", nPublic=int(j["nPublic"]), IC=[g1_from_json(p) for p in j["IC"]],
                vk_alfa_1=g1_from_json(j["vk_alfa_1"]), vk_beta_2=g2_from_json(j["vk_beta_2"]),
                vk_gamma_2=g2_from_json(j["vk_gamma_2"]), vk_delta_2=g2_from_json(j["vk_delta_2"]))


def proof_to_json(proof):
    """websnark groth16GenProof result shape."""
    return dict(pi_a=g1_to_json(proof["pi_a"]), pi_b=g2_to_json(proof["pi_b"]),
                pi_c=g1_to_json(proof["pi_c"]), protocol="groth")


def proof_from_json(j):
    return dict(pi_a=g1_from_json(j["pi_a"]), pi_b=g2_from_json(j["pi_b"]),
                pi_c=g1_from_json(j["pi_c"]), protocol="groth")
