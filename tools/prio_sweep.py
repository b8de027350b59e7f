#!/usr/bin/env python3
"""Experiment: does giving one of the five proof chains a higher CUDA stream priority shorten the overlapped
proof?  One synthetic setup, then per configuration a fresh context (ZKR_STREAM_PRIO is read at zkr_ctx_create),
key load, and timed proofs with the witness resident (zkr_prove_dev).  python tools/prio_sweep.py [--shape tx_2p20]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_zk_rollups_b200 import _lib, keygen, prover, synth  # noqa: E402

TOXIC = (0x1234567890ABCDEF1234567890ABCDEF1234567, 0x2222222222222222222222222222222222221,
         0x3333333333333333333333333333333333333331, 0x44444444444444444444444444444444441,
         0x555555555555555555555555555555555555555551)
# name -> "p0,..,p5[;H][;Ln][;Cn]": stream priorities, H = NTT pipeline before the MSMs, Ln = boundary-level granularity,
# Cn = forced MSM window size.  Results of the round-1 sweeps: profiles/r01_sched_sweep.json, r01_window_sweep.json.
CONFIGS = {"equal": "0,0,0,0,0,0", "b2_highest_h_high": "-1,0,0,-2,0,0", "ab_high": "0,-1,-1,0,0,0",
           "b2_highest_ab_h_high": "-1,-1,-1,-2,0,0", "all_but_c_high": "-1,-1,-1,-1,0,0", "h_first": "0,0,0,0,0,0;H",
           "lvl4": "0,0,0,0,0,0;L4", "c16": "0,0,0,0,0,0;C16", "c19": "0,0,0,0,0,0;C19"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="tx_2p20")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--out", default="gpurun_out/prio_sweep.json")
    a = ap.parse_args()
    import torch
    nc, npub = synth.SHAPES[a.shape]
    r1, w = synth.generate(nc, npub, seed=11)
    gp = prover.Groth16Prover(0)
    pk_bin, _ = keygen.synth_setup(gp.ctx, r1, TOXIC)
    gp.close()
    wb = np.frombuffer(synth.witness_bytes(w), dtype=np.uint8)
    res = {}
    ref = None
    for name, prio in CONFIGS.items():
        os.environ["ZKR_STREAM_PRIO"] = prio.split(";")[0]
        os.environ["ZKR_H_FIRST"] = "1" if prio.endswith(";H") else "0"
        os.environ["ZKR_LEVEL_LOG_BIG"] = prio.split(";L")[1] if ";L" in prio else "0"
        os.environ["ZKR_MSM_C"] = prio.split(";C")[1] if ";C" in prio else "0"
        gp = prover.Groth16Prover(0)
        L = gp.L
        key = gp.load_key(pk_bin)
        stream = torch.cuda.current_stream()
        _lib.check(L.zkr_ctx_set_stream(gp.ctx, C.c_void_p(stream.cuda_stream)))
        w_dev = torch.from_numpy(wb.copy()).cuda()
        out = torch.zeros(256, dtype=torch.uint8, device="cuda")
        n = r1.nVars
        rb = np.frombuffer((0x1F2E3D4C5B6A79881122334455667788 << 64 | 0x99AABBCCDDEEFF00).to_bytes(32, "little"), dtype=np.uint8)
        sb = np.frombuffer((0x0123456789ABCDEF << 100 | 77).to_bytes(32, "little"), dtype=np.uint8)

        def run():
            _lib.check(L.zkr_prove_dev(gp.ctx, key, C.c_void_p(w_dev.data_ptr()), n, _lib.buf_ptr(rb), _lib.buf_ptr(sb),
                                       C.c_void_p(out.data_ptr())))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            run()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        pb = out.cpu().numpy().tobytes()
        ref = ref or pb
        res[name] = {"prio": prio, "prove_ms": round(ms, 3), "same_proof": pb == ref}
        print(name, res[name], flush=True)
        gp.close()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
