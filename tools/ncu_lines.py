"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export per source line: instructions, samples, SIMT width."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10: continue
    if r[0] != "" and r[0] != "0":   # source line row (aggregated over its SASS)
        try:
            ix = {h: i for i, h in enumerate(hdr)}
            # the first "Source" column is the cuda source; metrics columns follow after the 4th column
            inst = float(r[ix["Instructions Executed"]] or 0); tinst = float(r[ix["Thread Instructions Executed"]] or 0)
            smp = float(r[ix["# Samples"]] or 0)
            agg[(cur_file, int(r[0]))] = (inst, tinst, smp, r[1].strip()[:110])
        except Exception as e:
            pass
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[2] for v in agg.values())
print(f"total warp-inst {tot_i:.3e}  samples {tot_s:.0f}")
byfile = collections.Counter(); byfile_s = collections.Counter()
for (f, l), v in agg.items(): byfile[f] += v[0]; byfile_s[f] += v[2]
for f in byfile: print(f"  {f:20s} inst {100*byfile[f]/tot_i:5.1f}%  samples {100*byfile_s[f]/tot_s:5.1f}%")
print("--- top lines by samples")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{f:16s}:{l:4d} smp {100*v[2]/tot_s:5.2f}% inst {100*v[0]/tot_i:5.2f}% width {v[1]/max(v[0],1):5.1f} | {v[3]}")
