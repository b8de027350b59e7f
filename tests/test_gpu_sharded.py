"""GPU parity of the multi-GPU paths, emulated on ONE B200: `world` contexts on cuda:0, each with its own
stream, wired with zkr_comm_connect_local.  The kernels, the peer addressing and the flag barrier are
exactly the ones that run across GPUs (tools/multigpu_check.py does the same over CUDA IPC under torchrun);
results must be bit-identical to the single-GPU transforms / MSMs, which are pinned to the oracle."""
import ctypes as C
import random
import threading

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import groth16 as g
from simple_zk_rollups_b200 import _lib, sharding as sh
from helpers import pack, pack_g1, pack_g2, unpack

pytestmark = pytest.mark.gpu
R = bn.R


class Ranks:
    def __init__(self, world, max_elems):
        import torch
        self.L = _lib.lib()
        self.world = world
        self.ctxs, self.streams, self.comms = [], [], []
        for r in range(world):
            h = C.c_void_p()
            _lib.check(self.L.zkr_ctx_create(0, C.byref(h)))
            st = torch.cuda.Stream(device=0)
            _lib.check(self.L.zkr_ctx_set_stream(h, C.c_void_p(st.cuda_stream)))
            self.ctxs.append(h)
            self.streams.append(st)
            self.comms.append(sh.Comm(h, r, world, max_elems))
        sh.Comm.connect_local(self.comms)

    def prepare_ntt(self, log_n):
        """Build every rank's twiddle tables up front: on ONE device the cudaMalloc / stream sync inside the
        lazy table build could wait on a peer's spinning barrier kernel (emulation artefact, see test_sharded_msm)."""
        dummy = np.zeros((1 << log_n) * 32, dtype=np.uint8)
        for c in self.ctxs:
            _lib.check(self.L.zkr_ntt(c, _lib.buf_ptr(dummy), log_n, 0, 0))

    def each(self, fn):
        """run fn(rank) on one host thread per rank (calls that synchronise wait for their peers)"""
        out, errs = [None] * self.world, []

        def body(r):
            try:
                out[r] = fn(r)
            except Exception as e:      # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=body, args=(r,)) for r in range(self.world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        return out

    def sync(self):
        for c in self.ctxs:
            _lib.check(self.L.zkr_ctx_synchronize(c))
        for c in self.comms:
            c.check()

    def close(self):
        for c in self.comms:
            c.close()
        for c in self.ctxs:
            self.L.zkr_ctx_destroy(c)


@pytest.fixture(scope="module", params=[1, 2, 4])
def ranks(request):
    rk = Ranks(request.param, 1 << 19)
    yield rk
    rk.close()


def brev_rows(x, bits):
    idx = np.array([g.bit_reverse(i, bits) for i in range(1 << bits)])
    return np.ascontiguousarray(x[idx])


def single(zctx, x, log_n, mode):
    buf = x.copy().reshape(-1)
    _lib.check(_lib.lib().zkr_ntt(zctx, _lib.buf_ptr(buf), log_n, mode, 0))
    return buf.reshape(-1, 32)


def rand_fr(n, seed):
    rs = np.random.RandomState(seed)
    x = rs.randint(0, 256, size=(n, 32), dtype=np.uint8)
    x[:, 31] &= 0x1F
    return x




@pytest.mark.parametrize("log_n,k0", [(16, None), (19, None), (19, 5), (18, 4)])
def test_sharded_ntt_all_modes(zctx, ranks, log_n, k0, monkeypatch):
    """k0 = None: the default plan (one local pass after the exchange).  k0 forced small: rows longer than a
    tile, i.e. the short-pass + full-pass structure that sizes >= 2^23 use."""
    if k0 is not None:
        monkeypatch.setenv("ZKR_NTT_SHARD_K0", str(k0))
        assert sh.rows_log(log_n, ranks.world) == max(k0, sh.log2_exact(ranks.world))
    world = ranks.world
    n = 1 << log_n
    x = rand_fr(n, log_n)
    xb = brev_rows(x, log_n)
    ranks.prepare_ntt(log_n)
    for mode in (sh.NTT_FORWARD, sh.NTT_INVERSE, sh.NTT_COSET_FORWARD, sh.NTT_COSET_INVERSE):
        want = single(zctx, x, log_n, mode)                         # natural in, natural out
        # DIF: COLS/natural -> ROWS/bit-reversed
        for r, c in enumerate(ranks.comms):
            c.upload(0, sh.cols_slab(x, log_n, r, world))
        ranks.each(lambda r: ranks.comms[r].ntt(log_n, mode | sh.NTT_BITREV_OUT, 0))
        ranks.sync()
        got = sh.from_rows_slabs([c.download(1, n // world) for c in ranks.comms], log_n)
        assert np.array_equal(got, brev_rows(want, log_n)), "DIF mode %d world %d" % (mode, world)
        # DIT: ROWS/bit-reversed -> COLS/natural  (source buffer 1 this time)
        for r, c in enumerate(ranks.comms):
            c.upload(1, sh.rows_slab(xb, log_n, r, world))
        ranks.each(lambda r: ranks.comms[r].ntt(log_n, mode | sh.NTT_BITREV_IN, 1))
        ranks.sync()
        got = sh.from_cols_slabs([c.download(0, n // world) for c in ranks.comms], log_n)
        assert np.array_equal(got, want), "DIT mode %d world %d" % (mode, world)


def test_sharded_ntt_vs_oracle_and_chain(ranks):
    """2^14 against the recursive oracle, then the H-pipeline-style chain DIF^-1 -> DIT round trip without
    any re-layout in between."""
    world, log_n = ranks.world, 14
    ranks.prepare_ntt(log_n)
    n = 1 << log_n
    rng = random.Random(7)
    vals = [rng.randrange(R) for _ in range(n)]
    vals[0], vals[1] = R - 1, 0
    x = pack(vals).reshape(n, 32)
    for r, c in enumerate(ranks.comms):
        c.upload(0, sh.cols_slab(x, log_n, r, world))
    ranks.each(lambda r: ranks.comms[r].ntt(log_n, sh.NTT_INVERSE | sh.NTT_BITREV_OUT, 0))
    ranks.sync()
    got = sh.from_rows_slabs([c.download(1, n // world) for c in ranks.comms], log_n)
    want = g.ntt(vals, inverse=True)
    assert unpack(got.reshape(-1)) == [want[g.bit_reverse(i, log_n)] for i in range(n)]
    ranks.each(lambda r: ranks.comms[r].ntt(log_n, sh.NTT_FORWARD | sh.NTT_BITREV_IN, 1))
    ranks.sync()
    back = sh.from_cols_slabs([c.download(0, n // world) for c in ranks.comms], log_n)
    assert np.array_equal(back, x)


@pytest.mark.parametrize("group,n", [(1, 1), (1, 5), (1, 1000), (2, 333)])
def test_sharded_msm(zctx, ranks, group, n):
    L = _lib.lib()
    world = ranks.world
    rng = random.Random(31 * group + n + world)
    cur = bn.G1 if group == 1 else bn.G2
    gen = bn.G1_GEN if group == 1 else bn.G2_GEN
    base = [cur.mul(gen, rng.randrange(1, 1 << 40)) for _ in range(8)]
    pts = [cur.add(base[i % 8], base[(i * 5 + 3) % 8]) if i % 3 else cur.mul(base[i % 8], i + 2) for i in range(n)]
    if n > 3:
        pts[2] = None                                               # an infinity entry inside a slice
    arr = (pack_g1(pts) if group == 1 else pack_g2(pts)).reshape(n, -1)
    scal = [rng.randrange(R) for _ in range(n)]
    if n > 4:
        scal[3], scal[4] = 0, R - 1
    sc = pack(scal).reshape(n, 32)
    ob = 64 if group == 1 else 128
    full = C.c_void_p()
    _lib.check(L.zkr_bases_load(zctx, group, _lib.buf_ptr(arr), n, 0, C.byref(full)))
    want = np.zeros(ob, dtype=np.uint8)
    _lib.check(L.zkr_msm(zctx, full, _lib.buf_ptr(sc), n, 0, _lib.buf_ptr(want)))
    L.zkr_bases_free(full)
    ref = cur.to_affine(g.msm_naive(cur, pts, scal))
    flat = [0] * (ob // 32) if ref is None else ([ref[0], ref[1]] if group == 1 else [ref[0][0], ref[0][1], ref[1][0], ref[1][1]])
    assert unpack(want) == flat

    # phases are separated by joins: on ONE device a device-wide synchronising call (cudaFree inside
    # zkr_bases_load / zkr_bases_free) would wait for a peer's spinning barrier kernel -- an artefact of the
    # emulation, not of the multi-GPU path (one device per rank)
    slices = [sh.point_range(n, r, world) for r in range(world)]

    def load(r):
        lo, hi = slices[r]
        b, dk = C.c_void_p(), C.c_void_p()
        sl = np.ascontiguousarray(arr[lo:hi])
        _lib.check(L.zkr_bases_load(ranks.ctxs[r], group, _lib.buf_ptr(sl) if hi > lo else None, hi - lo, 0, C.byref(b)))
        _lib.check(L.zkr_dev_malloc(ranks.ctxs[r], 32 * (hi - lo) + 32, C.byref(dk)))
        if hi > lo:
            _lib.check(L.zkr_dev_upload(ranks.ctxs[r], dk, _lib.buf_ptr(np.ascontiguousarray(sc[lo:hi])), 32 * (hi - lo)))
        return b, dk
    loaded = ranks.each(load)

    def run(r):
        lo, hi = slices[r]
        b, dk = loaded[r]
        out = ranks.comms[r].msm(b, dk.value, hi - lo, on_device=True)
        out2 = ranks.comms[r].msm(b, dk.value, hi - lo, on_device=True)   # second call uses the other slot parity
        assert np.array_equal(out, out2)
        return out[:ob]
    outs = ranks.each(run)

    def free(r):
        L.zkr_bases_free(loaded[r][0])
        _lib.check(L.zkr_dev_free(ranks.ctxs[r], loaded[r][1]))
    ranks.each(free)
    for got in outs:
        assert np.array_equal(got, want)


@pytest.mark.parametrize("n_constraints,n_public", [(300, 4), (1500, 9)])
def test_sharded_prove(ranks, n_constraints, n_public):
    """One proof split over `world` ranks (point-range-sharded MSMs, partials gathered through peer memory)
    is byte-identical to the oracle's proof -- hence to the single-GPU proof."""
    from oracle import binfmt as bf
    from simple_zk_rollups_b200 import keygen, synth
    L = _lib.lib()
    world = ranks.world
    toxic = (1234567891011, 222222222222223, 3333333333333331, 44444444444447, 5555555555555557)
    r1, w = synth.generate(n_constraints, n_public, seed=11 + n_constraints)
    pk_bin, vk = keygen.synth_setup(ranks.ctxs[0], r1, toxic)
    pk_o, vk_o, _ = g.setup(r1.to_dicts(), toxic)
    r, s = 0x1234567890ABCDEF1122334455667788, R - 5
    want, pub = g.gen_proof(pk_o, w, r, s)
    want = g.proof_to_bytes(want)
    wit = np.frombuffer(bf.binarify_witness(w), dtype=np.uint8)
    rb = np.frombuffer(int(r).to_bytes(32, "little"), dtype=np.uint8)
    sb = np.frombuffer(int(s).to_bytes(32, "little"), dtype=np.uint8)
    pkb = np.frombuffer(pk_bin, dtype=np.uint8) if isinstance(pk_bin, (bytes, bytearray)) else pk_bin

    def load(rk):
        h = C.c_void_p()
        _lib.check(L.zkr_pkey_load_bin_sharded(ranks.ctxs[rk], _lib.buf_ptr(pkb), pkb.size, rk, world, C.byref(h)))
        return h
    keys = ranks.each(load)
    # world 4: stages back to back on each rank's own stream (4 x 6 concurrently busy streams would share
    # hardware queues with the spinning barrier kernels on one device); world <= 2 keeps the 5-stream overlap
    for c in ranks.ctxs:
        _lib.check(L.zkr_ctx_set_serial(c, 1 if world > 2 else 0))

    def prove(rk):
        outs = []
        for _ in range(2):
            out = np.zeros(256, dtype=np.uint8)
            st = _lib.Stats()
            _lib.check(L.zkr_prove_sharded(ranks.comms[rk].h, keys[rk], _lib.buf_ptr(wit), wit.size // 32,
                                           _lib.buf_ptr(rb), _lib.buf_ptr(sb), _lib.buf_ptr(out), C.byref(st)))
            outs.append(out.tobytes())
        return outs
    proofs = ranks.each(prove)
    for c in ranks.ctxs:
        _lib.check(L.zkr_ctx_set_serial(c, 0))
    ranks.each(lambda rk: L.zkr_pkey_free(keys[rk]))
    for outs in proofs:
        assert outs[0] == want and outs[1] == want
    assert g.verify(vk_o, g.proof_from_bytes(want), pub)
