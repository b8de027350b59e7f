"""Byte-level packing helpers shared by the GPU parity tests."""
import numpy as np

from oracle import bn254 as bn

MONT = 1 << 256


def pack(vals):
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype=np.uint8).copy()


def unpack(arr):
    b = arr.tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def pack_g1(pts):
    return pack([c * MONT % bn.Q for p in pts for c in ((0, 0) if p is None else p)])


def pack_g2(pts):
    flat = []
    for p in pts:
        flat += [0, 0, 0, 0] if p is None else [p[0][0], p[0][1], p[1][0], p[1][1]]
    return pack([c * MONT % bn.Q for c in flat])


def unpack_g1(arr):
    rinv = pow(MONT, -1, bn.Q)
    v = [x * rinv % bn.Q for x in unpack(arr)]
    return [None if (v[i] == 0 and v[i + 1] == 0) else (v[i], v[i + 1]) for i in range(0, len(v), 2)]


def unpack_g2(arr):
    rinv = pow(MONT, -1, bn.Q)
    v = [x * rinv % bn.Q for x in unpack(arr)]
    return [None if not any(v[i:i + 4]) else ((v[i], v[i + 1]), (v[i + 2], v[i + 3])) for i in range(0, len(v), 4)]
