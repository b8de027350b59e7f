#!/usr/bin/env python3
"""Multi-GPU parity + throughput of the sharded paths (BASELINE.json configs[2..3] at 2/4/8 GPUs), one
process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py [--ntt-logs 20,22,24] [--msm-logs 20,22] [--out gpurun_out/multigpu_N.json]

torch.distributed (NCCL) is plumbing only: it carries the 64-byte IPC handles, the barriers around the
timed regions and the max-over-ranks reduction.  The data path is libzkr's own kernels storing into peer
HBM over NVLink.  Checks (every rank, exact bytes):
  * sharded NTT == this rank's slab of the single-GPU transform of the same vector (computed on the same GPU);
  * sharded MSM == (sum_i k_i s_i mod r) * G computed on the host for points P_i = s_i * G."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from simple_zk_rollups_b200 import _lib, sharding as sh  # noqa: E402
from sweep import Q, W_model, dot_mod_r  # noqa: E402

MODMUL_PEAK = 67.9e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ntt-logs", default="20,22,24")
    ap.add_argument("--msm-logs", default="20,22")
    ap.add_argument("--prove-shapes", default="", help="comma list of synth.SHAPES to prove sharded over all ranks (e.g. tx_2p20)")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    ctx = C.c_void_p()
    _lib.check(L.zkr_ctx_create(local, C.byref(ctx)))
    stream = torch.cuda.current_stream()
    _lib.check(L.zkr_ctx_set_stream(ctx, C.c_void_p(stream.cuda_stream)))
    ntt_logs = [int(v) for v in args.ntt_logs.split(",") if v]
    msm_logs = [int(v) for v in args.msm_logs.split(",") if v]
    comm = sh.Comm(ctx, rank, world, (1 << max(ntt_logs + [16])) // world)
    if world > 1:
        comm.connect_torch()
    res = {"world": world, "ntt": [], "msm": [], "prove": []}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, reps):
        fn()
        best = 1e30
        for _ in range(reps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, max_over_ranks(e0.elapsed_time(e1)))
        return best

    # ------------------------------------------------------------------ NTT
    for lg in ntt_logs:
        n = 1 << lg
        nl = n // world
        rng = np.random.default_rng(1000 + lg)
        x = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        x[:, 31] &= 0x1F
        ref = torch.from_numpy(x.reshape(-1)).cuda()
        _lib.check(L.zkr_ntt(ctx, C.c_void_p(ref.data_ptr()), lg, sh.NTT_FORWARD | sh.NTT_BITREV_OUT, 1))
        comm.upload(0, sh.cols_slab(x, lg, rank, world))
        comm.ntt(lg, sh.NTT_FORWARD | sh.NTT_BITREV_OUT, 0)
        torch.cuda.synchronize()
        comm.check()
        got = comm.download(1, nl)
        want = ref.cpu().numpy().reshape(n, 32)[rank * nl:(rank + 1) * nl]
        ok_dif = bool(np.array_equal(got, want))
        # chain back: DIT inverse of the bit-reversed ROWS data returns the COLS slab of x
        comm.ntt(lg, sh.NTT_INVERSE | sh.NTT_BITREV_IN, 1)
        torch.cuda.synchronize()
        comm.check()
        ok_rt = bool(np.array_equal(comm.download(0, nl), sh.cols_slab(x, lg, rank, world)))
        del ref
        t_dif = timed(lambda: comm.ntt(lg, sh.NTT_FORWARD | sh.NTT_BITREV_OUT, 0), args.reps)
        t_dit = timed(lambda: comm.ntt(lg, sh.NTT_INVERSE | sh.NTT_BITREV_IN, 1), args.reps)
        comm.check()
        oks = torch.tensor([int(ok_dif and ok_rt)], device="cuda")
        if world > 1:
            dist.all_reduce(oks, op=dist.ReduceOp.MIN)
        row = dict(log_n=lg, world=world, dif_ms=round(t_dif, 4), dit_inverse_ms=round(t_dit, 4),
                   dif_gb_per_s=round(64.0 * n / (t_dif * 1e-3) / 1e9, 1),
                   dit_gb_per_s=round(64.0 * n / (t_dit * 1e-3) / 1e9, 1),
                   exchange_mb_per_rank=round(32.0 * nl * (world - 1) / world / 1e6, 2), correct=bool(oks.item()))
        res["ntt"].append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)

    # ------------------------------------------------------------------ MSM (G1)
    for lg in msm_logs:
        n = 1 << lg
        lo, hi = sh.point_range(n, rank, world)
        rng = np.random.default_rng(2000 + lg)
        s64 = rng.integers(1, 1 << 63, size=n, dtype=np.uint64)
        k = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        k[:, 31] &= 0x1F
        sel = rng.random(n) < 0.03                       # rollup-like: 3 % of the scalars are {0,1}
        k[sel] = 0
        k[sel, 0] = rng.integers(0, 2, size=int(sel.sum()), dtype=np.uint8)
        sc_pts = np.zeros((hi - lo, 32), dtype=np.uint8)
        sc_pts[:, :8] = s64[lo:hi].view(np.uint8).reshape(hi - lo, 8)
        pts = np.empty((hi - lo) * 64, dtype=np.uint8)
        _lib.check(L.zkr_synth_points(ctx, 1, _lib.buf_ptr(sc_pts), hi - lo, _lib.buf_ptr(pts)))
        bases = C.c_void_p()
        _lib.check(L.zkr_bases_load(ctx, 1, _lib.buf_ptr(pts), hi - lo, 0, C.byref(bases)))
        cc, ww = C.c_int(), C.c_int()
        _lib.check(L.zkr_bases_info(bases, None, C.byref(cc), C.byref(ww), None))
        d_k = torch.from_numpy(np.ascontiguousarray(k[lo:hi]).reshape(-1)).cuda()
        out = comm.msm(bases, d_k.data_ptr(), hi - lo, on_device=True)[:64]
        ok = None
        if rank == 0:
            e = dot_mod_r(k.reshape(-1), s64)
            esc = np.frombuffer(int(e).to_bytes(32, "little"), dtype=np.uint8).copy()
            exp_m = np.empty(64, dtype=np.uint8)
            _lib.check(L.zkr_synth_points(ctx, 1, _lib.buf_ptr(esc), 1, _lib.buf_ptr(exp_m)))
            rinv = pow(1 << 256, -1, Q)
            exp = b"".join((int.from_bytes(exp_m[i:i + 32].tobytes(), "little") * rinv % Q).to_bytes(32, "little")
                           for i in range(0, 64, 32))
            ok = exp == out.tobytes()
        same = torch.from_numpy(out.copy()).cuda()
        if world > 1:                                    # every rank must hold the same sum
            ref0 = same.clone()
            dist.broadcast(ref0, 0)
            agree = torch.tensor([int(torch.equal(ref0, same))], device="cuda")
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)
            agree = bool(agree.item())
        else:
            agree = True
        t = timed(lambda: comm.msm(bases, d_k.data_ptr(), hi - lo, on_device=True), args.reps)
        mm = W_model(n)
        row = dict(group=1, log_n=lg, world=world, c=cc.value, windows=ww.value, ms=round(t, 4),
                   gpts_per_s=round(n / t / 1e6, 4), frac_of_modmul_peak_all_gpus=round(mm / (t * 1e-3) / (MODMUL_PEAK * world), 4),
                   correct=ok, ranks_agree=agree)
        res["msm"].append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
        L.zkr_bases_free(bases)
        del d_k
    # ------------------------------------------------------------------ one proof split over all ranks
    for shape in [v for v in args.prove_shapes.split(",") if v]:
        from simple_zk_rollups_b200 import keygen, synth
        toxic = (0x1234567890ABCDEF1234567890ABCDEF1234567, 0x2222222222222222222222222222222222221,
                 0x3333333333333333333333333333333333333331, 0x44444444444444444444444444444444441,
                 0x555555555555555555555555555555555555555551)
        nc, npub = synth.SHAPES[shape]
        r1, w = synth.generate(nc, npub, seed=11)
        pk_bin, _ = keygen.synth_setup(ctx, r1, toxic)
        wit = np.frombuffer(synth.witness_bytes(w), dtype=np.uint8)
        rb = np.frombuffer(int(0x1F2E3D4C5B6A7988 << 64 | 5).to_bytes(32, "little"), dtype=np.uint8)
        sb = np.frombuffer(int(0x0123456789ABCDEF << 100 | 77).to_bytes(32, "little"), dtype=np.uint8)
        full, part = C.c_void_p(), C.c_void_p()
        _lib.check(L.zkr_pkey_load_bin(ctx, _lib.buf_ptr(pk_bin), pk_bin.size, C.byref(full)))
        want = np.zeros(256, dtype=np.uint8)
        st1 = _lib.Stats()

        def prove_single():
            _lib.check(L.zkr_prove(ctx, full, _lib.buf_ptr(wit), wit.size // 32, _lib.buf_ptr(rb), _lib.buf_ptr(sb),
                                   _lib.buf_ptr(want), C.byref(st1)))
        t_single = timed(prove_single, args.reps)
        L.zkr_pkey_free(full)
        _lib.check(L.zkr_pkey_load_bin_sharded(ctx, _lib.buf_ptr(pk_bin), pk_bin.size, rank, world, C.byref(part)))
        got = np.zeros(256, dtype=np.uint8)
        st2 = _lib.Stats()

        def prove_sharded():
            _lib.check(L.zkr_prove_sharded(comm.h, part, _lib.buf_ptr(wit), wit.size // 32, _lib.buf_ptr(rb),
                                           _lib.buf_ptr(sb), _lib.buf_ptr(got), C.byref(st2)))
        t_sh = timed(prove_sharded, args.reps)
        same = torch.tensor([int(np.array_equal(got, want))], device="cuda")
        if world > 1:
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
        L.zkr_pkey_free(part)
        row = dict(shape=shape, world=world, constraints=nc, single_gpu_ms=round(t_single, 3), sharded_ms=round(t_sh, 3),
                   speedup=round(t_single / t_sh, 3), identical_to_single_gpu_proof=bool(same.item()),
                   sharded_stage_ms={k: round(v, 3) for k, v in st2.as_dict().items() if k.endswith("_ms")})
        res["prove"].append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
    if rank == 0 and args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(res, open(args.out, "w"), indent=1)
    comm.close()
    L.zkr_ctx_destroy(ctx)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    bad = [r for r in res["ntt"] if not r["correct"]] + [r for r in res["msm"] if r["correct"] is False or not r["ranks_agree"]]
    bad += [r for r in res["prove"] if not r["identical_to_single_gpu_proof"]]
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
