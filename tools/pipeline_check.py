#!/usr/bin/env python3
"""Experiment: proofs/s on ONE GPU with K independent proofs in flight (K contexts on the same device, each
with its own stream set and key replica) against one at a time.  python tools/pipeline_check.py [--shape tx_2p20]"""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="tx_2p20")
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--out", default="gpurun_out/pipeline_check.json")
    a = ap.parse_args()
    import torch
    nc, npub = synth.SHAPES[a.shape]
    r1, w = synth.generate(nc, npub, seed=11)
    gp0 = prover.Groth16Prover(0)
    pk_bin, _ = keygen.synth_setup(gp0.ctx, r1, TOXIC)
    gp0.close()
    wb = np.frombuffer(synth.witness_bytes(w), dtype=np.uint8)
    n = r1.nVars
    res = {}
    ref = None
    for K in (1, 2, 3):
        gps, keys, streams, wd, outs = [], [], [], [], []
        for k in range(K):
            gp = prover.Groth16Prover(0)
            keys.append(gp.load_key(pk_bin))
            st = torch.cuda.Stream()
            _lib.check(gp.L.zkr_ctx_set_stream(gp.ctx, C.c_void_p(st.cuda_stream)))
            gps.append(gp)
            streams.append(st)
            wd.append(torch.from_numpy(wb.copy()).cuda())
            outs.append(torch.zeros(256, dtype=torch.uint8, device="cuda"))
        L = gps[0].L

        def run(i):
            k = i % K
            _lib.check(L.zkr_prove_dev(gps[k].ctx, keys[k], C.c_void_p(wd[k].data_ptr()), n, None, None,
                                       C.c_void_p(outs[k].data_ptr())))
        for i in range(3 * K):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main_s = torch.cuda.current_stream()
        e0.record(main_s)
        for st in streams:
            st.wait_event(e0)
        for i in range(a.steps * K):
            run(i)
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            main_s.wait_event(ev)
        e1.record(main_s)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        pb = outs[0].cpu().numpy().tobytes()
        ref = ref or pb
        same = all(o.cpu().numpy().tobytes() == ref for o in outs)
        res[K] = {"in_flight": K, "proofs": a.steps * K, "total_ms": round(ms, 2), "proofs_per_s": round(a.steps * K / ms * 1e3, 2),
                  "ms_per_proof": round(ms / (a.steps * K), 3), "same_proof": same}
        print(res[K], flush=True)
        for gp in gps:
            gp.close()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
