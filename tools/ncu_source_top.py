#!/usr/bin/env python3
"""Hot spots of `ncu -i x.ncu-rep --page source --csv --print-source sass`: per captured kernel, the SASS instructions
that collect the most warp-stall samples, with the dominant stall reason, plus totals per stall reason and per opcode.

  python tools/ncu_source_top.py <source.csv> [--top 25] [--kernel-index I]
"""
import csv
import sys
from collections import Counter


def sections(path):
    cur = None
    with open(path, newline="") as f:
        for row in csv.reader(f):
            if not row:
                continue
            if row[0] == "Kernel Name":
                cur = {"name": row[1], "hdr": None, "rows": []}
                yield_me = cur
                secs.append(yield_me)
            elif cur is not None and cur["hdr"] is None:
                cur["hdr"] = row
            elif cur is not None:
                cur["rows"].append(row)


secs = []


def main():
    path = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    only = int(sys.argv[sys.argv.index("--kernel-index") + 1]) if "--kernel-index" in sys.argv else None
    sections(path)
    for si, s in enumerate(secs):
        if only is not None and si != only:
            continue
        h = s["hdr"]
        ia, isrc, ismp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        tot = Counter()
        by_op = Counter()
        rows = []
        for r in s["rows"]:
            try:
                n = int(r[ismp] or 0)
            except ValueError:
                continue
            st = {c: int(r[i] or 0) for i, c in stall_cols}
            for c, v in st.items():
                tot[c] += v
            op = r[isrc].split()[0] if r[isrc].split() else "?"
            if op.startswith("@"):
                op = r[isrc].split()[1]
            by_op[op.split(".")[0]] += n
            rows.append((n, r[ia], r[isrc], st, int(r[iex] or 0)))
        total = sum(n for n, *_ in rows)
        print("## kernel %d: %s  (%d SASS instructions, %d samples)" % (si, s["name"][:70], len(rows), total))
        print("stall totals: " + ", ".join("%s %.1f%%" % (c[6:], 100.0 * v / max(1, sum(tot.values()))) for c, v in tot.most_common(9)))
        print("samples by opcode: " + ", ".join("%s %.1f%%" % (o, 100.0 * v / max(1, total)) for o, v in by_op.most_common(10)))
        for n, a, src, st, ex in sorted(rows, key=lambda x: -x[0])[:top]:
            dom = max(st.items(), key=lambda kv: kv[1])
            print("  %5.2f%%  %-8s %-60s %s=%d  exec=%d" % (100.0 * n / max(1, total), a[-6:], src[:60], dom[0][6:], dom[1], ex))
        print()


if __name__ == "__main__":
    main()
