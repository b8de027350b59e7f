"""Synthetic proving / verifying keys for rollup-shaped R1CS, made on the GPU (zkr_synth_setup).

Stands where the reference runs `snarkjs setup --protocol groth` (prover/package.json:34,37); the
toxic waste is an explicit input, so these are TEST / BENCHMARK keys.  Output is the websnark
binary proving key -- byte-identical to binarifyProvingKey(snarkjs pk JSON) for the same
(R1CS, toxic waste) -- plus the verifying key as Python ints (and as the binary block zkr_vkey_load_bin takes, vk["bin"]).
"""
import ctypes as C
import struct

import numpy as np

from . import _lib
from .binarify import Q, R

_RINV_Q = pow(1 << 256, -1, Q)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _pols_section(ptr, row, cid, pool_mont_u32, n):
    """writeTransformedPolynomial x n (binarify.ts:104-113,170-177), vectorised."""
    nnz = int(row.size)
    words = np.zeros(n + 9 * nnz, dtype=np.uint32)
    ptr64 = ptr.astype(np.int64)
    words[np.arange(n, dtype=np.int64) + 9 * ptr64[:-1]] = (ptr64[1:] - ptr64[:-1]).astype(np.uint32)
    if nnz:
        sig = np.repeat(np.arange(n, dtype=np.int64), (ptr64[1:] - ptr64[:-1]))
        base = sig + 1 + 9 * np.arange(nnz, dtype=np.int64)
        words[base] = row
        words[(base + 1)[:, None] + np.arange(8)] = pool_mont_u32[cid]
    return words.view(np.uint8)


def _g1(buf):
    x = int.from_bytes(buf[:32], "little") * _RINV_Q % Q
    y = int.from_bytes(buf[32:64], "little") * _RINV_Q % Q
    return None if x == 0 else (x, y)


def _g2(buf):
    v = [int.from_bytes(buf[32 * i:32 * i + 32], "little") * _RINV_Q % Q for i in range(4)]
    return None if v[0] == 0 and v[1] == 0 else ((v[0], v[1]), (v[2], v[3]))


def synth_setup(ctx, r1cs, toxic):
    """r1cs: simple_zk_rollups_b200.synth.R1CS; toxic = (tau, alpha, beta, gamma, delta) ints.
    -> (pk_bin: np.ndarray[uint8], vk: dict of int points)."""
    L = _lib.lib()
    n, l = r1cs.nVars, r1cs.nPublic
    bits, m = r1cs.domain()
    csc = r1cs.csc(with_inputs=True)
    keep = []
    desc = _lib.R1csCsc()
    desc.n_vars, desc.n_public, desc.n_constraints, desc.domain_size = n, l, r1cs.nConstraints, m
    desc.n_pool = len(r1cs.pool)
    for k in "abc":
        ptr, row, cid = (_u32(x) for x in csc[k.upper()])
        keep += [ptr, row, cid]
        setattr(desc, "ptr_" + k, ptr.ctypes.data)
        setattr(desc, "row_" + k, row.ctypes.data)
        setattr(desc, "cid_" + k, cid.ctypes.data)
    pool = np.frombuffer(b"".join(int(c).to_bytes(32, "little") for c in r1cs.pool), dtype=np.uint8).copy()
    desc.pool = pool.ctypes.data
    tox = np.frombuffer(b"".join(int(t % R).to_bytes(32, "little") for t in toxic), dtype=np.uint8).copy()
    out_a = np.empty(64 * n, dtype=np.uint8)
    out_b1 = np.empty(64 * n, dtype=np.uint8)
    out_b2 = np.empty(128 * n, dtype=np.uint8)
    out_c = np.empty(64 * (n - l - 1), dtype=np.uint8)
    out_h = np.empty(64 * m, dtype=np.uint8)
    out_vk = np.empty(192 + 384 + 64 * (l + 1), dtype=np.uint8)
    _lib.check(L.zkr_synth_setup(ctx, C.byref(desc), _lib.buf_ptr(tox), *[_lib.buf_ptr(x) for x in
                                 (out_a, out_b1, out_b2, out_c, out_h, out_vk)]))
    pool_mont = np.frombuffer(b"".join(((int(c) << 256) % R).to_bytes(32, "little") for c in r1cs.pool),
                              dtype=np.uint32).reshape(-1, 8)
    secA = _pols_section(*[_u32(x) for x in csc["A"]], pool_mont, n)
    secB = _pols_section(*[_u32(x) for x in csc["B"]], pool_mont, n)
    vkb = out_vk.tobytes()
    fixed = vkb[0:192] + vkb[192:320] + vkb[448:576]          # alfa1 beta1 delta1 | beta2 | delta2
    ptrs, off = [], 40 + len(fixed)
    for sec in (secA, secB, out_a, out_b1, out_b2, out_c, out_h):
        ptrs.append(off)
        off += sec.size
    if off >= 1 << 32:
        raise ValueError("websnark binary key offsets are u32 (binarify.ts:155-161): key too large")
    head = struct.pack("<10I", n, l, m, *ptrs)
    pk_bin = np.concatenate([np.frombuffer(head + fixed, dtype=np.uint8), secA, secB, out_a, out_b1, out_b2,
                             out_c, out_h])
    assert pk_bin.size == off
    vk = dict(protocol="groth", nPublic=l, vk_alfa_1=_g1(vkb[0:64]), vk_beta_2=_g2(vkb[192:320]),
              vk_gamma_2=_g2(vkb[320:448]), vk_delta_2=_g2(vkb[448:576]),
              IC=[_g1(vkb[576 + 64 * i:640 + 64 * i]) for i in range(l + 1)],
              bin=vkb)                 # the block zkr_vkey_load_bin takes
    return pk_bin, vk
