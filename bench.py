#!/usr/bin/env python3
"""Benchmark of the Groth16 prove hot path (BASELINE.json metric: "BN254 Groth16 prove ms + proofs/s; ...").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shape tx_2p20]

One "step" = one complete proof (witness -> 256-byte proof) of the rollup-shaped circuit named in
config.workload.  Default workload = BASELINE.json configs[1]: the rollup circuit scaled to ~2^20
constraints (shape of BatchProcessTx(16, 6): 858 400 constraints, 577 public inputs, m = 2^20),
synthetic R1CS / witness / key (see simple_zk_rollups_b200/synth.py, keygen.py).
  value  proofs/s with the witness already resident in HBM (zkr_prove_dev), CUDA events on the launch stream
  e2e    proofs/s through the host API (zkr_prove_batch; zkr_prove per call is reported beside it): pinned-host
         witness H2D + proof D2H inside the timed region
N > 1 (torchrun, one rank per GPU): independent proofs, one per GPU, no data-path collective ("weak").
--impl reference: the C restatement of the reference's CPU algorithm (oracle/c, kind "port": the reference's
own prover is un-vendored JavaScript/WASM that cannot run in this image) on all host cores, rank 0 only.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOXIC = (0x1234567890ABCDEF1234567890ABCDEF1234567, 0x2222222222222222222222222222222222221,
         0x3333333333333333333333333333333333333331, 0x44444444444444444444444444444444441,
         0x555555555555555555555555555555555555555551)
BASELINE_CONFIG = {"tx": "BASELINE.json configs[0]", "tx_2p20": "BASELINE.json configs[1]",
                   "tx_2p22": "BASELINE.json configs[4] (per-GPU unit of the throughput batch)"}
MODMUL_IMAD = 136            # 8x8-limb CIOS: 128 wide MACs + 8
MADD_MODMULS = 10            # XYZZ mixed add 8M + 2S


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(shape, seed):
    from simple_zk_rollups_b200 import synth
    nc, npub = synth.SHAPES[shape]
    t0 = time.time()
    r1, w = synth.generate(nc, npub, seed=seed)
    log("[bench] synthetic R1CS %s: %d constraints, %d public, nVars %d, nnz %s (%.1f s)" % (
        shape, nc, npub, r1.nVars, r1.nnz(), time.time() - t0))
    return r1, w


def run_ours(args):
    import numpy as np
    import torch
    from simple_zk_rollups_b200 import _lib, keygen, prover, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun --nproc-per-node %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libzkr has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    gp = prover.Groth16Prover(local)
    stream = torch.cuda.current_stream()
    _lib.check(L.zkr_ctx_set_stream(gp.ctx, C.c_void_p(stream.cuda_stream)))

    r1, w = make_workload(args.shape, seed=11 + rank)       # distinct witnesses per rank, same shape
    t0 = time.time()
    pk_bin, vk = keygen.synth_setup(gp.ctx, r1, TOXIC)
    key = gp.load_key(pk_bin)
    info = gp.key_info(key)
    log("[bench] key: %.1f MB binary, %.2f GB resident in HBM (%.1f s setup+load)" % (
        pk_bin.size / 1e6, info["device_bytes"] / 1e9, time.time() - t0))
    n, m = info["nVars"], info["domainSize"]
    wbytes = synth.witness_bytes(w)
    w_host = torch.frombuffer(bytearray(wbytes), dtype=torch.uint8).pin_memory()
    w_dev = w_host.cuda()
    proof_dev = torch.zeros(256, dtype=torch.uint8, device="cuda")
    rs = (0x1F2E3D4C5B6A79881122334455667788 << 64 | 0x99AABBCCDDEEFF00, 0x0123456789ABCDEF << 100 | 77)
    rb = np.frombuffer(int(rs[0]).to_bytes(32, "little"), dtype=np.uint8)
    sb = np.frombuffer(int(rs[1]).to_bytes(32, "little"), dtype=np.uint8)

    def prove_dev():
        _lib.check(L.zkr_prove_dev(gp.ctx, key, C.c_void_p(w_dev.data_ptr()), n, _lib.buf_ptr(rb), _lib.buf_ptr(sb),
                                   C.c_void_p(proof_dev.data_ptr())))

    out = np.zeros(256, dtype=np.uint8)
    stats = _lib.Stats()

    def prove_host():
        _lib.check(L.zkr_prove(gp.ctx, key, C.c_void_p(w_host.data_ptr()), n, _lib.buf_ptr(rb), _lib.buf_ptr(sb),
                               _lib.buf_ptr(out), C.byref(stats)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = gp.kernel_launches()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = gp.kernel_launches() - l0
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    # parity guard: the proof must be what the host API returns for the same inputs
    prove_dev()
    torch.cuda.synchronize()
    prove_host()
    assert bytes(proof_dev.cpu().numpy().tobytes()) == out.tobytes(), "device and host API proofs differ"

    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = timed(prove_dev, args.steps, args.warmup)
    clocks = sampler.stop()
    ms_e2e, _ = timed(prove_host, args.steps, max(args.warmup, 1))

    # e2e throughput through the batch entry point (the shape of many genTxVerifierProof calls): host witness
    # buffers, every proof's H2D and D2H inside the timed region, witness i+1 uploaded while proof i runs
    ctxs1, pks1 = (C.c_void_p * 1)(gp.ctx), (C.c_void_p * 1)(key)
    w_host2 = w_host.clone().pin_memory()
    rs_pair = np.concatenate([rb, sb])

    def prove_batch_host(nproofs):
        wptrs = (C.c_void_p * nproofs)(*[(w_host if i % 2 == 0 else w_host2).data_ptr() for i in range(nproofs)])
        rsb = np.tile(rs_pair, nproofs)
        outb = np.zeros(256 * nproofs, dtype=np.uint8)
        _lib.check(L.zkr_prove_batch(ctxs1, pks1, 1, wptrs, n, nproofs, _lib.buf_ptr(rsb), _lib.buf_ptr(outb)))
        return outb

    ob = prove_batch_host(max(args.warmup, 2))
    assert ob[-256:].tobytes() == out.tobytes(), "batch and single-call proofs differ"
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ob = prove_batch_host(args.steps)
    e1.record(stream)
    barrier()
    ms_batch = e0.elapsed_time(e1)
    assert ob[-256:].tobytes() == out.tobytes() and ob[:256].tobytes() == out.tobytes()
    if world > 1:
        t = torch.tensor([ms_batch], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_batch = float(t.item())
    stage_ms = {k: round(v, 3) for k, v in stats.as_dict().items() if k.endswith("_ms")}

    result = None
    if rank == 0:
        # ---- roofline of the dominant kernel (G1 bucket accumulation), measured live on a serialised pass
        _lib.check(L.zkr_ctx_set_serial(gp.ctx, 1))
        _lib.check(L.zkr_ctx_set_profile(gp.ctx, 1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        prof_steps = min(args.steps, 4)
        e0.record(stream)
        for _ in range(prof_steps):
            prove_dev()
        e1.record(stream)
        torch.cuda.synchronize()
        serial_ms = e0.elapsed_time(e1) / prof_steps
        prof = {}
        for pid, name in ((0, "accum_g1"), (1, "accum_g2"), (2, "ntt_pass")):
            tot, cnt, units = C.c_double(), C.c_int(), C.c_double()
            _lib.check(L.zkr_ctx_profile_read(gp.ctx, pid, C.byref(tot), C.byref(cnt), C.byref(units)))
            prof[name] = (tot.value, cnt.value, units.value)
        _lib.check(L.zkr_ctx_set_profile(gp.ctx, 0))
        _lib.check(L.zkr_ctx_set_serial(gp.ctx, 0))
        imad_peak, ms_mb = C.c_double(), C.c_float()
        _lib.check(L.zkr_microbench(gp.ctx, 0, 100000, C.byref(imad_peak), C.byref(ms_mb)))
        modmul_peak = C.c_double()
        _lib.check(L.zkr_microbench(gp.ctx, 2, 20000, C.byref(modmul_peak), C.byref(ms_mb)))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass
        g1_ms, g1_cnt, g1_madds = prof["accum_g1"]
        roofline = None
        if g1_cnt:
            imads = g1_madds * MADD_MODMULS * MODMUL_IMAD
            ach = imads / (g1_ms * 1e-3) / 1e12
            roofline = {"kernel": "k_accum_affine<Fq> (G1 bucket accumulation, XYZZ mixed adds)",
                        "bound": "imad", "achieved": round(ach, 3), "peak": round(imad_peak.value / 1e12, 3),
                        "unit": "TIMAD/s", "frac": round(ach / (imad_peak.value / 1e12), 4),
                        "traffic": traffic.get("accum_g1_bytes_per_launch"),
                        "peak_source": "plain IMAD chain measured in this run (zkr_microbench); MEASURED_PEAKS.json has no integer peak",
                        "practical_peak_frac": round((g1_madds * MADD_MODMULS / (g1_ms * 1e-3)) / modmul_peak.value, 4),
                        "practical_peak_note": "vs the register-resident Fq modmul chain measured in this run (%.1f G modmul/s): "
                                               "IMAD.WIDE issues at half the IMAD rate" % (modmul_peak.value / 1e9),
                        "launches": g1_cnt, "avg_launch_ms": round(g1_ms / g1_cnt, 4),
                        "algorithmic_units_per_launch": "%.0f mixed adds x 10 modmul x 136 IMAD" % (g1_madds / g1_cnt),
                        "share_of_serial_step": round(g1_ms / prof_steps / serial_ms, 4),
                        "measured": "serialised pass of the same step (zkr_ctx_set_serial), CUDA events on the launch stream"}
        nt_ms, nt_cnt, nt_elems = prof["ntt_pass"]
        roofline_ntt = None
        if nt_cnt:
            gbs = 64.0 * nt_elems / (nt_ms * 1e-3) / 1e9
            roofline_ntt = {"kernel": "k_ntt_pass (one <=11-bit radix pass over 2^%d Fr elements)" % (m.bit_length() - 1),
                            "bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak, "unit": "GB/s",
                            "frac": round(gbs / hbm_peak, 4), "traffic": traffic.get("ntt_pass_bytes_per_launch"),
                            "peak_source": hbm_src, "launches": nt_cnt, "avg_launch_ms": round(nt_ms / nt_cnt, 4),
                            "note": "64 B per element per pass (read once + write once); the pass is IMAD-bound "
                                    "(~6 modmul per 64 B), see DESIGN.md"}
            # the binding roof: butterfly modmuls, (m/2) log2 m per transform, 6 transforms per proof
            log_m = m.bit_length() - 1
            bfly = 6.0 * prof_steps * (m // 2) * log_m
            roofline_ntt["imad"] = {
                "bound": "imad", "butterfly_modmuls_per_proof": int(bfly / prof_steps),
                "achieved_timad_per_s": round(bfly * MODMUL_IMAD / (nt_ms * 1e-3) / 1e12, 3),
                "frac_of_plain_imad_peak": round(bfly * MODMUL_IMAD / (nt_ms * 1e-3) / imad_peak.value, 4),
                "frac_of_practical_modmul_peak": round(bfly / (nt_ms * 1e-3) / modmul_peak.value, 4),
                "note": "algorithmic butterflies only; inter-pass / coset twiddle products (~1 extra modmul per element "
                        "per pass) are not counted"}
        # ---- CPU baseline: the C restatement of the reference algorithm on this box's host cores
        cpu = None if args.no_cpu else cpu_baseline(pk_bin, wbytes, rs, out.tobytes(), args)
        value = world * args.steps / (ms * 1e-3)
        e2e_val = world * args.steps / (ms_batch * 1e-3)
        result = {
            "metric": "groth16_proofs_per_s", "value": round(value, 3), "unit": "proofs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (8x32-bit-limb Montgomery, integer pipe)",
            "data": "synthetic",
            "config": {"workload": "BN254 Groth16 prove, rollup-shaped circuit %s: %d constraints, %d public, nVars %d, "
                                   "domain 2^%d; fixed (r,s); key resident in HBM" % (
                                       args.shape, r1.nConstraints, r1.nPublic, n, m.bit_length() - 1),
                       "baseline_config": BASELINE_CONFIG.get(args.shape, "BASELINE.json configs[1] shape family"),
                       "parallelism": "one independent proof per GPU",
                       "l2": "per-proof working set (%.1f GB of window tables + sort buffers) >> 126 MB L2; no flush needed" % (
                           info["device_bytes"] / 1e9)},
            "prove_ms": round(ms / args.steps, 4), "prove_ms_e2e": round(ms_e2e / args.steps, 4),
            "prove_ms_serial": round(serial_ms, 4),
            "e2e": {"value": round(e2e_val, 3), "unit": "proofs/s", "h2d_bytes_per_step": 32 * n + 64,
                    "d2h_bytes_per_step": 256, "ms_per_step": round(ms_batch / args.steps, 4),
                    "api": "zkr_prove_batch on host witness buffers (pinned): per proof H2D of the witness + (r,s), "
                           "D2H of the 256-byte proof and the range flags; witness i+1 is uploaded while proof i runs",
                    "single_call": {"api": "zkr_prove (one blocking call per proof, nothing overlapped)",
                                    "value": round(world * args.steps / (ms_e2e * 1e-3), 3),
                                    "ms_per_step": round(ms_e2e / args.steps, 4)}},
            "gpu_launches": int(launches), "clocks": clocks, "stage_ms_overlapped": stage_ms,
            "roofline": roofline, "roofline_ntt": roofline_ntt, "cpu_baseline": cpu,
        }
    gp.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if result is not None:
        print(json.dumps(result), flush=True)


def cpu_baseline(pk_bin, wbytes, rs, gpu_proof, args):
    """oracle/c (kind 'port') on all host cores; one full proof of the same workload, checked against the GPU's."""
    from oracle import cbind
    cores = os.cpu_count() or 1
    t0 = time.time()
    got = cbind.prove(pk_bin, wbytes, rs[0], rs[1], mode=1, threads=cores)
    dt = time.time() - t0
    ok = got == gpu_proof
    log("[bench] cpu_baseline: %.2f s per proof on %d threads; proof %s the GPU's" % (dt, cores, "==" if ok else "!="))
    return {"value": round(1.0 / dt, 5), "unit": "proofs/s", "cores": cores, "kind": "port",
            "sample": "1 full proof of the same workload (same key, witness, r, s); Pippenger + iterative NTT, "
                      "websnark calcH structure; proof bytes %s the GPU proof" % ("equal" if ok else "DIFFER FROM"),
            "seconds_per_proof": round(dt, 3), "matches_gpu_proof": ok}


def run_reference(args):
    """Reference arm: the CPU port of the reference's algorithm, all host threads, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import numpy as np
    from oracle import cbind
    from simple_zk_rollups_b200 import synth
    cbind.build()
    cores = os.cpu_count() or 1
    r1, w = make_workload(args.shape, seed=11)
    wbytes = synth.witness_bytes(w)
    key_path = os.path.join(ROOT, "gpurun_out", "_ref_key_%s.bin" % args.shape)
    pk_bin = None
    try:
        # the key is workload data, not code under test: made once with the GPU setup when a GPU is there
        import torch
        if torch.cuda.is_available():
            from simple_zk_rollups_b200 import keygen, prover
            gp = prover.Groth16Prover(int(os.environ.get("LOCAL_RANK", "0")))
            pk_bin, _ = keygen.synth_setup(gp.ctx, r1, TOXIC)
            gp.close()
    except Exception as e:          # noqa: BLE001
        log("[bench] GPU key generation unavailable (%s)" % e)
    if pk_bin is None and os.path.exists(key_path):
        pk_bin = np.fromfile(key_path, dtype=np.uint8)
    if pk_bin is None:
        print(json.dumps({"impl": "reference", "unavailable": "no proving key for the workload (needs a GPU to run the synthetic setup)"}))
        return
    rs = (0x1F2E3D4C5B6A79881122334455667788 << 64 | 0x99AABBCCDDEEFF00, 0x0123456789ABCDEF << 100 | 77)
    mode = args.ref_mode
    if args.ref_threads:
        cores = args.ref_threads
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.time()
        cbind.prove(pk_bin, wbytes, rs[0], rs[1], mode=mode, threads=cores)
        dt = time.time() - t0
        if i >= args.warmup:
            times.append(dt)
        log("[bench] reference step %d: %.2f s" % (i, dt))
    total = sum(times)
    val = len(times) / total
    n = r1.nVars
    bits, m = r1.domain()
    print(json.dumps({
        "impl": "reference", "metric": "groth16_proofs_per_s", "value": round(val, 5), "unit": "proofs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total / len(times), 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (4x64-bit-limb Montgomery)",
        "data": "synthetic",
        "config": {"workload": "BN254 Groth16 prove, rollup-shaped circuit %s: %d constraints, %d public, nVars %d, "
                               "domain 2^%d; fixed (r,s)" % (args.shape, r1.nConstraints, r1.nPublic, n, bits),
                   "baseline_config": BASELINE_CONFIG.get(args.shape, "BASELINE.json configs[1] shape family")},
        "cpu_baseline": {"value": round(val, 5), "unit": "proofs/s", "cores": cores, "kind": "port",
                         "sample": "every step is 1 full proof of the workload (C restatement of websnark groth16GenProof: "
                                   + ("Pippenger multiexp + iterative NTT" if mode == 1 else
                                      "snarkjs arithmetic structure: one double-and-add multiplication per point, recursive radix-2 FFT")
                                   + ", %d host thread(s))" % cores},
        "e2e": {"value": round(val, 5), "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="tx_2p20", help="withdraw | tx | tx_2p20 | tx_2p22 (simple_zk_rollups_b200.synth.SHAPES)")
    ap.add_argument("--ref-mode", type=int, default=1, choices=[0, 1],
                    help="--impl reference: 1 = Pippenger + iterative NTT (default), 0 = snarkjs arithmetic structure (SURVEY 8(d) CPU-A)")
    ap.add_argument("--ref-threads", type=int, default=0, help="--impl reference: host threads (default: all)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs under ncu)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] note: timing rules ask for >= 3 warm-up steps")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
