"""Per-source-line instruction / stall-sample table from `ncu --page source --csv --print-source cuda,sass`.

usage: python tools/ncu_lines.py SRC.csv N_WINDOWS [file-substring ...]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
nwin = float(sys.argv[2]); want = sys.argv[3:]
secs = []; cur = None
for r in rows:
    if r and r[0] == "File Path": cur = {'file': r[1], 'rows': []}; secs.append(cur)
    elif r and r[0] == "Line No": cur['hdr'] = r
    elif cur is not None and 'hdr' in cur and r: cur['rows'].append(r)
grand = 0; gs = 0
for s in secs:
    h = s['hdr']; ii = h.index("Instructions Executed"); isamp = h.index("# Samples")
    stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    agg = []
    for r in s['rows']:
        if r[2] != "-" or not r[0]: continue
        try: n = int(r[ii]); sm = int(r[isamp])
        except ValueError: continue
        if n > 0 or sm > 0:
            top = sorted(((int(r[i]) if r[i].isdigit() else 0, h[i][6:]) for i in stall), reverse=True)[:2]
            agg.append((n, sm, int(r[0]), r[1].strip()[:90], " ".join("%s:%d" % (b, a) for a, b in top if a)))
    t = sum(a[0] for a in agg); ss = sum(a[1] for a in agg); grand += t; gs += ss
    print("== %s: inst/window %.1f samples %d" % (s['file'].split('/')[-1], t / nwin, ss))
    if any(w in s['file'] for w in want):
        for a in sorted(agg, key=lambda x: x[2]):
            if a[0] / nwin >= 2 or a[1] >= 40: print("%7.1f %5d L%-4d %-90s %s" % (a[0] / nwin, a[1], a[2], a[3], a[4]))
print("total inst/window %.1f samples %d" % (grand / nwin, gs))
