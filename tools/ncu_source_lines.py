"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump by (kernel, file:line):
python tools/ncu_source_lines.py dump.csv [kernel substring] [top N]"""
import csv, sys, collections
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
func = None; hdr = None; fpath = None
agg = collections.defaultdict(lambda: [0, 0, 0])   # inst, samples, source text
text = {}
kern_tot = collections.Counter()
mode = None
for row in csv.reader(open(path, errors="ignore")):
    if not row: continue
    if row[0] == "File Path": fpath = row[1]; continue
    if row[0] == "Function Name": func = row[1]; hdr = None; continue
    if row[0] == "Line No" or row[0] == "Address":
        hdr = row; continue
    if hdr is None or func is None or want not in func: continue
    if not row[0].strip().isdigit(): continue            # SASS rows (empty line number) repeat the counts of their source line
    d = {}
    for k, v in zip(hdr, row):
        d.setdefault(k, v)
    if "Instructions Executed" not in d: continue
    try:
        ie = int(d["Instructions Executed"] or 0); sm = int(d.get("# Samples") or 0)
    except ValueError:
        continue
    # rows of the CUDA-source view have "Line No"; keep only those (the SASS view repeats the counts)
    key = (func.split("(")[0][-40:], (fpath or "").split("/")[-1], d["Line No"])
    agg[key][0] += ie; agg[key][1] += sm; text[key] = d["Source"].strip()[:110]
    kern_tot[key[0]] += ie
by = 1 if (len(sys.argv) > 4 and sys.argv[4] == "samples") else 0      # 5th argument "samples": order by stall samples instead
smp_tot = sum(v[1] for v in agg.values())
for k, (ie, sm, _) in sorted(agg.items(), key=lambda kv: -kv[1][by])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%s  %s" % (100.0 * ie / max(kern_tot[k[0]], 1), 100.0 * sm / max(smp_tot, 1), k[1], k[2], text[k]))
