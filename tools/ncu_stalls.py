"""Per source line: which stall reason the samples of an `ncu --page source --csv --print-source cuda,sass` export fall under.
Usage: ncu_stalls.py export.csv [reason=long_sb] [top=25]"""
import csv, sys, collections
path = sys.argv[1]; reason = sys.argv[2] if len(sys.argv) > 2 else "long_sb"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None
agg = collections.OrderedDict(); tot = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr) or r[0] in ("", "0", "Function Name"): continue
    try:
        v = float(r[ix["stall_" + reason]] or 0); smp = float(r[ix["# Samples"]] or 0)
    except ValueError:
        continue
    agg[(cur_file, int(r[0]))] = (v, smp, r[1].strip()[:120])
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            try: tot[h] += float(r[ix[h]] or 0)
            except ValueError: pass
all_s = sum(tot.values())
print("stall totals:", ", ".join(f"{k[6:]} {100*v/all_s:.1f}%" for k, v in tot.most_common(9)))
tr = sum(v[0] for v in agg.values())
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f:16s}:{l:4d} {reason} {100*v[0]/max(tr,1):5.1f}% (of all samples {100*v[0]/all_s:4.1f}%) | {v[2]}")
