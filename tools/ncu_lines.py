"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
cur_file = None; kern = None; hdr = None; last_line = '?'; last_src = ''
agg = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, ""]))
tot = collections.Counter(); totsamp = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": kern = r[1].split("(")[0].split("::")[-1] + (r[1].split("(")[0].split(",")[-1] if "," in r[1].split("(")[0] else ""); continue
    if r[0] == "Line No": hdr = r; i_inst = hdr.index("Instructions Executed"); i_samp = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr): continue
    try:
        inst = int(r[i_inst]); samp = int(r[i_samp])
    except ValueError:
        continue
    if r[0] != "":
        last_line = r[0]; last_src = r[1].strip()[:110]
        continue  # per-line summary row: the SASS rows that follow carry the counts
    key = (cur_file, last_line)
    a = agg[kern][key]; a[0] += inst; a[1] += samp; a[2] = last_src
    tot[kern] += inst; totsamp[kern] += samp
for k in agg:
    print("=====", k, "instructions", tot[k], "samples", totsamp[k])
    for (f, ln), (inst, samp, src) in sorted(agg[k].items(), key=lambda kv: -kv[1][1])[:topn]:
        print("%5.1f%% inst %5.1f%% samp  %s:%s  %s" % (100.0 * inst / max(1, tot[k]), 100.0 * samp / max(1, totsamp[k]), f, ln, src))
