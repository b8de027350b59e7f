#!/usr/bin/env python3
"""Turn ncu exports into the committed summaries under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.md> [--proof-launches N]
      launch list of `ncu --metrics gpu__time_duration.sum --clock-control none --csv ... bench.py`:
      per-kernel share of ONE proof (the last N launches of the file; N auto-detected from k_prep_scalars).
  python tools/ncu_summary.py raw <raw.csv> [<raw2.csv> ...] <out.md> [--traffic-json profiles/ncu_traffic.json]
      `ncu -i x.ncu-rep --page raw --csv` of --set full captures: the metrics the design argues with.
"""
import csv
import json
import re
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    # where the warps wait (per issue-active cycle); added for the NTT pass / G2 accumulation captures
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "launch__shared_mem_per_block_dynamic", "sm__maximum_warps_per_active_cycle_pct",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("<unnamed>::", "").replace("zkr::", "")
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("Fp<FqParams>", "Fq").replace("Fp<FrParams>", "Fr")
    return name[:90]


def launches(path, out, n_proof=None):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = [(short(r[ki]), float(r[vi].replace(",", "")) / 1e3) for r in rows]
    starts = [i for i, (k, _) in enumerate(seq) if k.startswith("k_prep_scalars")]
    if not starts:
        raise SystemExit("no k_prep_scalars launch in the list")
    # one proof = from the last k_prep_scalars to the end or the next one (profiling passes follow in bench.py)
    proofs = [(s, starts[i + 1] if i + 1 < len(starts) else len(seq)) for i, s in enumerate(starts)]
    lo, hi = proofs[1] if len(proofs) > 1 else proofs[0]        # second proof: warm tables
    if n_proof:
        hi = lo + n_proof
    one = seq[lo:hi]
    # cut at k_finish (end of the proof)
    for i, (k, _) in enumerate(one):
        if k.startswith("k_finish"):
            one = one[:i + 1]
            break
    agg = OrderedDict()
    for k, us in one:
        a = agg.setdefault(k, [0.0, 0])
        a[0] += us
        a[1] += 1
    tot = sum(a[0] for a in agg.values())
    with open(out, "w") as f:
        f.write("# Launch list of one proof (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n")
        f.write("Source: `%s` (`python bench.py --steps 2 --warmup 1 --no-cpu` under ncu on a B200; serialised, cold cache:\n"
                "compare SHARES, not absolutes).  One proof of the tx_2p20 workload = %d launches, %.2f ms summed.\n\n" % (
                    path, len(one), tot / 1e3))
        f.write("| share | time (us) | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write("| %.2f %% | %.1f | %d | `%s` |\n" % (100 * us / tot, us, n, k))
    print("wrote", out, "launches", len(one), "total ms", round(tot / 1e3, 2))


def raw(paths, out, traffic_json=None):
    sections = []
    traffic = {}
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units, rows = rows[0], rows[1], rows[2:]
        ki = hdr.index("Kernel Name")
        col = {}
        for i, h in enumerate(hdr):
            for k in KEEP:
                if h == k or h.endswith("." + k):
                    col.setdefault(k, i)
        for r in rows:
            name = short(r[ki])
            vals = OrderedDict()
            for k in KEEP:
                if k in col and r[col[k]] != "":
                    vals[k] = (r[col[k]], units[col[k]])
            sections.append((name, path, vals))
    with open(out, "w") as f:
        f.write("# ncu --set full captures (B200, `bench.py --steps 1 --warmup 1 --no-cpu`, workload tx_2p20)\n\n"
                "Numbers under ncu are cold-cache / serialised and are not bench values.  Exported with\n"
                "`ncu -i <rep> --page raw --csv`; this file keeps the metrics DESIGN.md argues with.\n")
        for name, path, vals in sections:
            f.write("\n### %s  (%s)\n" % (name, path.split("/")[-1]))
            for k, (v, u) in vals.items():
                f.write("    %-72s %s %s\n" % (k, v, u))
            try:
                rd = float(vals["dram__bytes_read.sum"][0].replace(",", ""))
                wr = float(vals["dram__bytes_write.sum"][0].replace(",", ""))
                mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                tb = rd * mul.get(vals["dram__bytes_read.sum"][1], 1.0) + wr * mul.get(vals["dram__bytes_write.sum"][1], 1.0)
                f.write("    dram read+write per launch: %.1f MB\n" % (tb / 1e6))
                traffic.setdefault(name, []).append(tb)
            except Exception:
                pass
    if traffic_json:
        t = {}
        try:
            t = json.load(open(traffic_json))
        except Exception:
            pass
        for name, v in traffic.items():
            avg = sum(v) / len(v)
            if name.startswith("k_accum_affine<Fq2"):
                t["accum_g2_bytes_per_launch"] = int(avg)
            elif name.startswith("k_accum_affine<Fq"):
                t["accum_g1_bytes_per_launch"] = int(avg)
            elif name.startswith("k_ntt_pass"):
                t["ntt_pass_bytes_per_launch"] = int(avg)
        t["_source"] = "%s (ncu --set full --clock-control none, B200; average over the captured launches)" % out
        json.dump(t, open(traffic_json, "w"), indent=1)
    print("wrote", out, len(sections), "kernels")


if __name__ == "__main__":
    a = sys.argv[1:]
    if a and a[0] == "launches":
        n = int(a[a.index("--proof-launches") + 1]) if "--proof-launches" in a else None
        launches(a[1], a[2], n)
    elif a and a[0] == "raw":
        tj = a[a.index("--traffic-json") + 1] if "--traffic-json" in a else None
        rest = [x for x in a[1:] if not x.startswith("--") and x != tj]
        raw(rest[:-1], rest[-1], tj)
    else:
        raise SystemExit(__doc__)
