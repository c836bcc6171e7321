#!/usr/bin/env python3
"""Per-kernel share of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv).  Usage: tools/launch_summary.py launches.csv [top]"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", "")) / 1000.0
    except ValueError:
        continue
    n = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    agg[n][0] += 1; agg[n][1] += v; tot += v
print(f"total {tot:.1f} us over {sum(c for c, _ in agg.values())} launches (cold-cache serialised times: the SHARE is what is comparable)")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{c:6d} launches {t:10.1f} us {100 * t / tot:5.1f} %  avg {t / c:7.1f} us  {n[:100]}")
