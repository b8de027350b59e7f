"""Throughput of the native snarkjs-JSON key ingestion (zkr_pkey_json_to_bin, host-only) against the Python
host mirror (json.loads + binarifyProvingKey), on a fabricated key of rollup shape (random field values: the
converter does no curve math).  Prints one JSON line.   python tools/bench_keyjson.py [--vars 100000]"""
import argparse
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simple_zk_rollups_b200 import binarify  # noqa: E402


def fabricate(n, n_public, seed=1):
    rnd = random.Random(seed)
    m = 1 << (n - 1).bit_length()
    fq = lambda: str(rnd.randrange(binarify.Q))
    g1 = lambda: [fq(), fq(), "1"]
    g2 = lambda: [[fq(), fq()], [fq(), fq()], ["1", "0"]]
    pol = lambda: {str(k): str(rnd.randrange(binarify.R)) for k in sorted(rnd.sample(range(m), rnd.choice((1, 2, 3))))}
    return dict(protocol="groth", nVars=n, nPublic=n_public, domainBits=m.bit_length() - 1, domainSize=m,
                polsA=[pol() for _ in range(n)], polsB=[pol() for _ in range(n)], polsC=[pol() for _ in range(n)],
                A=[g1() for _ in range(n)], B1=[g1() for _ in range(n)], B2=[g2() for _ in range(n)],
                C=[None] * (n_public + 1) + [g1() for _ in range(n - n_public - 1)], hExps=[g1() for _ in range(m)],
                vk_alfa_1=g1(), vk_beta_1=g1(), vk_delta_1=g1(), vk_beta_2=g2(), vk_delta_2=g2())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vars", type=int, default=100000)
    a = ap.parse_args()
    text = json.dumps(fabricate(a.vars, 73))
    t0 = time.perf_counter()
    native = binarify.binarifyProvingKeyJson(text)
    t1 = time.perf_counter()
    mirror = binarify.binarifyProvingKey(json.loads(text))
    t2 = time.perf_counter()
    assert native == mirror
    print(json.dumps({"n_vars": a.vars, "json_mb": round(len(text) / 1e6, 1), "bin_mb": round(len(native) / 1e6, 1),
                      "native_s": round(t1 - t0, 3), "native_json_mb_per_s": round(len(text) / 1e6 / (t1 - t0), 1),
                      "python_mirror_s": round(t2 - t1, 3), "speedup": round((t2 - t1) / (t1 - t0), 1),
                      "identical_bytes": True}))


if __name__ == "__main__":
    main()
