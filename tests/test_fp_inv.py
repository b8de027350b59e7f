"""The variable-time inversion used by k_finish (csrc/fp_inv.cuh) is plain C++: build it for the host and check it
against Python big-integer arithmetic (oracle moduli), including the edge values and the invalid inputs."""
import ctypes as C
import os
import random
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import bn254 as bn

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def inv_lib():
    d = tempfile.mkdtemp(prefix="zkr_fpinv_")
    so = os.path.join(d, "fp_inv_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "csrc", "fp_inv_host.cpp")])
    return C.CDLL(so)


def limbs(xs):
    return np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in xs), dtype=np.uint32).copy()


def ints(arr):
    b = arr.tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


@pytest.mark.parametrize("field,p", [(0, bn.Q), (1, bn.R)])
def test_binary_inverse_matches_python(inv_lib, field, p):
    m = np.zeros(8, dtype=np.uint32)
    inv_lib.fp_inv_modulus(field, m.ctypes.data_as(C.c_void_p))
    assert ints(m) == [p]
    rng = random.Random(20261017 + field)
    xs = [1, 2, 3, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, 1 << 255 % p, (1 << 253), (1 << 32) - 1, 1 << 32,
          pow(2, 256, p), pow(2, 512, p)]
    xs += [rng.randrange(1, p) for _ in range(3000)]
    xs += [rng.randrange(1, 1 << b) for b in (8, 31, 33, 64, 65, 127, 129, 200) for _ in range(20)]
    xs = [x % p or 1 for x in xs]
    a = limbs(xs)
    out = np.zeros_like(a)
    inv_lib.fp_inv_batch(field, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), len(xs))
    got = ints(out)
    for x, g in zip(xs, got):
        assert g == pow(x, -1, p), hex(x)


@pytest.mark.parametrize("field,p", [(0, bn.Q), (1, bn.R)])
def test_binary_inverse_invalid_inputs_terminate(inv_lib, field, p):
    """0 has no inverse (k_finish never passes it: infinity is handled before) and a >= p is out of contract:
    both must return 0 rather than loop."""
    xs = [0, p, 2 * p, 3 * p]
    a = limbs(xs)
    out = np.ones_like(a)
    inv_lib.fp_inv_batch(field, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), len(xs))
    got = ints(out)
    assert got[0] == 0 and got[1] == 0
