"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/ncu_launches.py file.csv [max_rows]"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 10]
hdr = rows[0]
iK, iG, iB, iV, iM = (hdr.index(k) for k in ("Kernel Name", "Grid Size", "Block Size", "Metric Value", "Metric Name"))
iU = hdr.index("Metric Unit")
tot = {}
lines = []
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    us = v / 1e3 if r[iU] in ("ns", "nsecond") else (v if r[iU] in ("us", "usecond") else v * 1e3 if r[iU] in ("ms", "msecond") else v)
    name = r[iK].split("(")[0][-60:]
    lines.append((name, r[iG], r[iB], us))
    tot[name] = tot.get(name, (0, 0.0))
    tot[name] = (tot[name][0] + 1, tot[name][1] + us)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for l in lines[:n]:
    print("%-62s %-18s %-12s %10.1f us" % l)
print("---- totals")
all_us = sum(v[1] for v in tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-62s x%-5d %12.1f us  %5.1f%%" % (k, v[0], v[1], 100 * v[1] / all_us))
print("total %.1f us" % all_us)
