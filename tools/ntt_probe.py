#!/usr/bin/env python3
"""NTT-only driver: a few DIF / DIT transforms of one size through the C-ABI (zkr_ntt, data resident in HBM).

  python tools/ntt_probe.py --log-n 20 --reps 20          timing per variant (CUDA events, JSON lines)
  ncu --set full -k regex:k_ntt_pass ... python tools/ntt_probe.py --log-n 20 --reps 1 --no-time

Kept apart from bench.py so that an ncu capture of the pass kernels does not have to replay a key setup.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_zk_rollups_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-time", action="store_true")
    args = ap.parse_args()
    import torch   # device buffers + events (plumbing)
    L = _lib.lib()
    ctx = C.c_void_p()
    _lib.check(L.zkr_ctx_create(0, C.byref(ctx)))
    stream = torch.cuda.current_stream()
    _lib.check(L.zkr_ctx_set_stream(ctx, C.c_void_p(stream.cuda_stream)))
    n = 1 << args.log_n
    rng = np.random.default_rng(7)
    x = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    x[:, 31] &= 0x1F
    d = torch.from_numpy(x.reshape(-1)).cuda()
    variants = (("forward_dif", 0 | 0x10), ("inverse_dif", 1 | 0x10), ("forward_dit", 0 | 0x20), ("inverse_dit", 1 | 0x20),
                ("coset_forward_dif", 2 | 0x10), ("coset_inverse_dit", 3 | 0x20))
    for name, mode in variants:
        def run():
            _lib.check(L.zkr_ntt(ctx, C.c_void_p(d.data_ptr()), args.log_n, mode, 1))
        run()
        torch.cuda.synchronize()
        if args.no_time:
            continue
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run()
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        best = min(ts)
        print(json.dumps(dict(log_n=args.log_n, variant=name, ms_best=round(best, 4), ms_median=round(float(np.median(ts)), 4),
                              gb_per_s=round(64.0 * n / best / 1e6, 1))), flush=True)
    L.zkr_ctx_destroy(ctx)


if __name__ == "__main__":
    main()
