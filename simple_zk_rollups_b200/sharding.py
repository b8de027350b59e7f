"""Host side of the multi-GPU paths (one process or one context per GPU): communicator wiring and the
slab layouts of the sharded four-step NTT / range-sharded MSM.  Pure index logic + ctypes calls into
libzkr; torch.distributed (NCCL or gloo) is used only as plumbing to exchange the 64-byte IPC handles.

Replaces the web-worker fan-out of websnark's multiexp / fft (inside groth16GenProof,
operator/src/snarks/common.ts:29) by a fan-out over the GPUs of one NVSwitch box.

Layouts (N = 2^log_n = R rows x C columns, R = 2^k0, k0 = rows_log(log_n, world), index = i*C + j):
    COLS slab of rank p : all rows, columns [p*C/world, (p+1)*C/world)   -- natural-order data
    ROWS slab of rank q : the contiguous slice [q*N/world, (q+1)*N/world) -- bit-reversed-order data
A DIF transform maps COLS/natural -> ROWS/bit-reversed, a DIT transform ROWS/bit-reversed -> COLS/natural,
each with ONE all-to-all that is fused into a kernel's write-back (remote stores over NVLink).
"""
import ctypes as C

import numpy as np

from . import _lib

NTT_FORWARD, NTT_INVERSE, NTT_COSET_FORWARD, NTT_COSET_INVERSE = 0, 1, 2, 3
NTT_BITREV_OUT, NTT_BITREV_IN = 0x10, 0x20
IPC_HANDLE_BYTES = 64


def log2_exact(x):
    lg = x.bit_length() - 1
    if x < 1 or (1 << lg) != x:
        raise ValueError("%d is not a power of two" % x)
    return lg


def rows_log_default(log_n, world):
    """k0 of the sharded four-step as csrc/ntt.cu (ntt_sharded_k0) chooses it by default; tests assert the
    two agree.  rows_log() asks the library, which also honours the ZKR_NTT_SHARD_K0 override."""
    g = log2_exact(world)
    k0 = 9 if log_n >= 23 else (log_n - 11 if log_n >= 14 else log_n // 2)
    return max(k0, g, 3)


def rows_log(log_n, world):
    return int(_lib.lib().zkr_ntt_sharded_rows_log(log_n, world))


def point_range(n, rank, world):
    """Contiguous point range [lo, hi) of `rank` for an MSM over n points (balanced to within one point)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def cols_slab(x, log_n, rank, world):
    """x: (N, 32) uint8 natural order -> rank's COLS slab, (N/world, 32)."""
    k0 = rows_log(log_n, world)
    R, Cc = 1 << k0, 1 << (log_n - k0)
    cl = Cc // world
    return np.ascontiguousarray(x.reshape(R, world, cl, 32)[:, rank]).reshape(R * cl, 32)


def rows_slab(x, log_n, rank, world):
    n = 1 << log_n
    return np.ascontiguousarray(x.reshape(n, 32)[rank * n // world:(rank + 1) * n // world])


def from_cols_slabs(slabs, log_n):
    world = len(slabs)
    k0 = rows_log(log_n, world)
    R, Cc = 1 << k0, 1 << (log_n - k0)
    cl = Cc // world
    out = np.empty((R, world, cl, 32), dtype=np.uint8)
    for p, s in enumerate(slabs):
        out[:, p] = s.reshape(R, cl, 32)
    return out.reshape(1 << log_n, 32)


def from_rows_slabs(slabs, log_n):
    return np.concatenate([s.reshape(-1, 32) for s in slabs], axis=0)


class Comm:
    """One rank's zkr_comm."""

    def __init__(self, ctx, rank, world, max_elems_per_rank):
        self.L = _lib.lib()
        self.ctx, self.rank, self.world = ctx, rank, world
        self.h = C.c_void_p()
        _lib.check(self.L.zkr_comm_create(ctx, rank, world, max_elems_per_rank, C.byref(self.h)))

    def export(self):
        buf = (C.c_char * IPC_HANDLE_BYTES)()
        _lib.check(self.L.zkr_comm_export(self.h, buf))
        return bytes(buf)

    def connect(self, handles):
        """handles: list of `world` 64-byte handles in rank order (cross-process, CUDA IPC)."""
        if len(handles) != self.world or any(len(h) != IPC_HANDLE_BYTES for h in handles):
            raise ValueError("need %d handles of %d bytes" % (self.world, IPC_HANDLE_BYTES))
        blob = b"".join(handles)
        _lib.check(self.L.zkr_comm_connect(self.h, blob))

    @staticmethod
    def connect_local(comms):
        """All ranks live in this process (one ctx per GPU, or several ctxs on one GPU for tests)."""
        arr = (C.c_void_p * len(comms))(*[c.h for c in comms])
        _lib.check(comms[0].L.zkr_comm_connect_local(arr, len(comms)))

    def connect_torch(self, group=None):
        """Exchange the IPC handles through torch.distributed (any backend) and map the peers."""
        import torch.distributed as dist
        handles = [None] * self.world
        dist.all_gather_object(handles, self.export(), group=group)
        self.connect(handles)
        dist.barrier(group=group)

    def buffer(self, which):
        return int(self.L.zkr_comm_buffer(self.h, which))

    def upload(self, which, arr):
        _lib.check(self.L.zkr_dev_upload(self.ctx, C.c_void_p(self.buffer(which)), _lib.buf_ptr(arr), arr.nbytes))

    def download(self, which, n_elems):
        out = np.empty((n_elems, 32), dtype=np.uint8)
        _lib.check(self.L.zkr_dev_download(self.ctx, _lib.buf_ptr(out), C.c_void_p(self.buffer(which)), out.nbytes))
        return out

    def barrier(self):
        _lib.check(self.L.zkr_comm_barrier(self.h))

    def check(self):
        _lib.check(self.L.zkr_comm_check(self.h))

    def ntt(self, log_n, mode, src_buf):
        """Sharded transform of exchange buffer src_buf -> buffer 1 - src_buf (asynchronous)."""
        _lib.check(self.L.zkr_ntt_sharded(self.h, log_n, mode, src_buf))
        return 1 - src_buf

    def msm(self, bases, scalars, n_local, on_device=False):
        """Range-sharded MSM: `bases` = this rank's slice.  -> affine std-form bytes (64 / 128), same on all ranks."""
        out = np.zeros(128, dtype=np.uint8)
        _lib.check(self.L.zkr_msm_sharded(self.h, bases, _lib.buf_ptr(scalars), n_local, 1 if on_device else 0,
                                          _lib.buf_ptr(out)))
        return out

    def close(self):
        if self.h:
            self.L.zkr_comm_destroy(self.h)
            self.h = C.c_void_p()
