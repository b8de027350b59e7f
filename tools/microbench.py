#!/usr/bin/env python3
"""Measure the integer-pipe peaks on the GPU box: IMAD, IMAD.WIDE,
register-resident Fq modmul chain, XYZZ mixed-add chain.  Writes gpurun_out/microbench.json."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_zk_rollups_b200 import _lib

L = _lib.lib()
h = C.c_void_p()
_lib.check(L.zkr_ctx_create(0, C.byref(h)))
res = {}
for name, which, iters in (("imad_per_s", 0, 100000), ("imad_wide_per_s", 1, 100000),
                           ("fq_modmul_per_s", 2, 20000), ("g1_madd_per_s", 3, 3000),
                           ("dfma_per_s", 4, 100000), ("dfma_imadw_pairs_per_s", 5, 100000),
                           ("dfma_iadd64_pairs_per_s", 6, 100000), ("iadd64_per_s", 7, 100000),
                           ("fq_inverse_fermat_per_s", 8, 40), ("fq_inverse_euclid_full_grid_per_s", 9, 40),
                           ("g1_batch_affine_add_b16_per_s", 10, 60), ("g1_batch_affine_add_b64_per_s", 11, 30),
                           ("g1_madd_lazy_per_s", 12, 3000), ("g2_madd_8warps_per_s", 13, 1000),
                           ("g2_madd_lazy_8warps_per_s", 14, 1000), ("g1_madd_lazy_sqr_per_s", 15, 3000)):
    ops, ms = C.c_double(), C.c_float()
    _lib.check(L.zkr_microbench(h, which, iters, C.byref(ops), C.byref(ms)))
    res[name] = ops.value
    res[name.replace("_per_s", "_ms")] = ms.value
res["modmul_imad_equiv_per_s"] = res["fq_modmul_per_s"] * 136
res["madd_modmul_equiv_per_s"] = res["g1_madd_per_s"] * 10
# batched affine: what one inversion costs in modmuls, and additions/s of the scheme against XYZZ mixed additions
res["fermat_inverse_in_modmuls"] = res["fq_modmul_per_s"] / res["fq_inverse_fermat_per_s"]
res["euclid_inverse_in_modmuls_full_grid"] = res["fq_modmul_per_s"] / res["fq_inverse_euclid_full_grid_per_s"]
res["batch_affine_b16_vs_madd"] = res["g1_batch_affine_add_b16_per_s"] / res["g1_madd_per_s"]
res["batch_affine_b64_vs_madd"] = res["g1_batch_affine_add_b64_per_s"] / res["g1_madd_per_s"]
res["g1_madd_lazy_sqr_vs_madd"] = res["g1_madd_lazy_sqr_per_s"] / res["g1_madd_per_s"]
res["g1_madd_lazy_vs_madd"] = res["g1_madd_lazy_per_s"] / res["g1_madd_per_s"]
res["g2_madd_lazy_vs_madd"] = res["g2_madd_lazy_8warps_per_s"] / res["g2_madd_8warps_per_s"]
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/microbench.json", "w"), indent=1)
print(json.dumps(res))
L.zkr_ctx_destroy(h)
