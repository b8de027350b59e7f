#!/usr/bin/env python3
"""Standalone sweeps of BASELINE.json configs[2..3] on one B200:
   G1 / G2 MSM (Gpts/s) and Fr NTT (GB/s) for 2^16 .. 2^26, kernel-only with inputs resident in HBM.
   python tools/sweep.py [--max-log 24] [--out gpurun_out/sweep.json]
Every MSM result is checked: points are s_i*G with 64-bit s_i, so sum k_i P_i = (sum k_i s_i mod r) G,
computed on the host with exact 16-bit-limb dot products and one fixed-base multiplication on the GPU."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_zk_rollups_b200 import _lib  # noqa: E402

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
MODMUL_PEAK = 67.9e9


def W_model(n):
    best = None
    for c in range(4, 24):
        w = -(-255 // c)
        cost = w * 10.0 * n + 28.0 * (1 << (c - 1))
        best = cost if best is None else min(best, cost)
    return best


def dot_mod_r(k_bytes, s_u64):
    """sum k_i * s_i mod r, exact (k: n x 32 B LE, s: n uint64)."""
    n = s_u64.size
    k16 = k_bytes.view(np.uint16).reshape(n, 16).astype(np.uint64)
    s16 = s_u64.view(np.uint16).reshape(n, 4).astype(np.uint64)
    tot = 0
    for a in range(16):
        ka = k16[:, a]
        for b in range(4):
            # products < 2^32; split the sum so that partial sums stay < 2^63
            acc = 0
            for lo in range(0, n, 1 << 26):
                acc += int(np.dot(ka[lo:lo + (1 << 26)], s16[lo:lo + (1 << 26), b]))
            tot += acc << (16 * (a + b))
    return tot % R


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log", type=int, default=24)
    ap.add_argument("--g2-max-log", type=int, default=22)
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--min-log", type=int, default=16)
    ap.add_argument("--g2-min-log", type=int, default=16)
    ap.add_argument("--skip-ntt", action="store_true")
    ap.add_argument("--g1-off", action="store_true", help="skip the G1 rows (G2-only runs)")
    args = ap.parse_args()
    if args.max_log > 26 or args.g2_max_log > 24:
        raise SystemExit("sweep sizes are capped at 2^26 (G1) / 2^24 (G2): beyond that the tables do not fit one GPU")
    L = _lib.lib()
    ctx = C.c_void_p()
    _lib.check(L.zkr_ctx_create(0, C.byref(ctx)))
    import torch   # device buffers + events (plumbing)
    stream = torch.cuda.current_stream()
    _lib.check(L.zkr_ctx_set_stream(ctx, C.c_void_p(stream.cuda_stream)))
    res = {"msm": [], "ntt": [], "modmul_peak_per_s": MODMUL_PEAK}
    rng = np.random.default_rng(2026)

    def timeit(fn, reps):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return min(ts), float(np.median(ts))

    for group, max_log in ((1, args.max_log), (2, min(args.g2_max_log, args.max_log))):
        if group == 1 and args.g1_off:
            continue
        for lg in range(args.min_log if group == 1 else args.g2_min_log, max_log + 1, 2):
            n = 1 << lg
            t0 = time.time()
            s64 = rng.integers(1, 1 << 63, size=n, dtype=np.uint64)
            sc_pts = np.zeros((n, 32), dtype=np.uint8)
            sc_pts[:, :8] = s64.view(np.uint8).reshape(n, 8)
            ab = 64 if group == 1 else 128
            pts = np.empty(n * ab, dtype=np.uint8)
            _lib.check(L.zkr_synth_points(ctx, group, _lib.buf_ptr(sc_pts), n, _lib.buf_ptr(pts)))
            bases = C.c_void_p()
            _lib.check(L.zkr_bases_load(ctx, group, _lib.buf_ptr(pts), n, 0, C.byref(bases)))
            npts, cc, ww, nbytes = C.c_uint64(), C.c_int(), C.c_int(), C.c_uint64()
            _lib.check(L.zkr_bases_info(bases, C.byref(npts), C.byref(cc), C.byref(ww), C.byref(nbytes)))
            t_load = time.time() - t0
            for dist in ("uniform", "rollup_like"):
                k = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
                k[:, 31] &= 0x1F                     # < 2^253 < r
                if dist == "rollup_like":            # 3 % of the scalars are {0,1}
                    sel = rng.random(n) < 0.03
                    k[sel] = 0
                    k[sel, 0] = rng.integers(0, 2, size=int(sel.sum()), dtype=np.uint8)
                d_k = torch.from_numpy(k.reshape(-1)).cuda()
                d_out = torch.zeros(256, dtype=torch.uint8, device="cuda")

                def run():
                    _lib.check(L.zkr_msm_dev(ctx, bases, C.c_void_p(d_k.data_ptr()), n, C.c_void_p(d_out.data_ptr())))
                best, med = timeit(run, args.reps)
                # correctness
                out = np.zeros(ab, dtype=np.uint8)
                _lib.check(L.zkr_msm(ctx, bases, C.c_void_p(d_k.data_ptr()), n, 1, _lib.buf_ptr(out)))
                e = dot_mod_r(k.reshape(-1), s64)
                esc = np.frombuffer(int(e).to_bytes(32, "little"), dtype=np.uint8).copy()
                exp_m = np.empty(ab, dtype=np.uint8)
                _lib.check(L.zkr_synth_points(ctx, group, _lib.buf_ptr(esc), 1, _lib.buf_ptr(exp_m)))
                rinv = pow(1 << 256, -1, Q)
                exp = b"".join((int.from_bytes(exp_m[i:i + 32].tobytes(), "little") * rinv % Q).to_bytes(32, "little")
                               for i in range(0, ab, 32))
                ok = exp == out.tobytes()
                mm = W_model(n) * (1 if group == 1 else 3)
                row = dict(group=group, log_n=lg, dist=dist, c=cc.value, windows=ww.value, ms_best=round(best, 4),
                           ms_median=round(med, 4), gpts_per_s=round(n / best / 1e6, 4),
                           modmul_model=mm, frac_of_modmul_peak=round(mm / (best * 1e-3) / MODMUL_PEAK, 4),
                           imad_frac_plain=round(mm * 136 / (best * 1e-3) / 18.57e12, 4),
                           table_gb=round(nbytes.value / 1e9, 3), load_s=round(t_load, 2), correct=bool(ok))
                res["msm"].append(row)
                print(json.dumps(row), flush=True)
                del d_k
            L.zkr_bases_free(bases)
            del pts
    for lg in range(16, (0 if args.skip_ntt else min(args.max_log + 2, 26)) + 1, 2):
        n = 1 << lg
        x = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        x[:, 31] &= 0x1F
        d = torch.from_numpy(x.reshape(-1)).cuda()
        for name, mode in (("dif_forward", 0 | 0x10), ("inverse_dit", 1 | 0x20), ("coset_forward_dif", 2 | 0x10)):
            def run():
                _lib.check(L.zkr_ntt(ctx, C.c_void_p(d.data_ptr()), lg, mode, 1))
            best, med = timeit(run, args.reps)
            mm = (n // 2) * lg
            row = dict(log_n=lg, op=name, ms_best=round(best, 4), ms_median=round(med, 4),
                       gb_per_s=round(64.0 * n / (best * 1e-3) / 1e9, 1), hbm_frac=round(64.0 * n / (best * 1e-3) / 6460.5e9, 4),
                       butterfly_modmul_frac_of_peak=round(mm / (best * 1e-3) / MODMUL_PEAK, 4))
            res["ntt"].append(row)
            print(json.dumps(row), flush=True)
        # round trip property (on fresh data: the timing loops transformed d in place)
        d.copy_(torch.from_numpy(x.reshape(-1)))
        _lib.check(L.zkr_ntt(ctx, C.c_void_p(d.data_ptr()), lg, 0, 1))
        _lib.check(L.zkr_ntt(ctx, C.c_void_p(d.data_ptr()), lg, 1, 1))
        assert np.array_equal(d.cpu().numpy().reshape(n, 32), x), "NTT round trip failed at 2^%d" % lg
        del d
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    L.zkr_ctx_destroy(ctx)


if __name__ == "__main__":
    main()
