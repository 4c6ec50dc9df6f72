"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_summary.py LAUNCHES.csv [kernel-substring-to-list-individually]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == 'ID':
        h = r; start = i + 2; break
ki = h.index('Kernel Name'); vi = h.index('Metric Value'); gi = h.index('Grid Size')
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= vi: continue
    n = re.sub(r'<.*', '', r[ki]); n = re.sub(r'\(.*', '', n).replace('void ', '')
    us = float(r[vi].replace(',', '')) / 1e3
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += us
    if len(sys.argv) > 2 and sys.argv[2] in n: print("   %s grid %s %.1f us" % (n, r[gi], us))
tot = sum(a[1] for a in agg.values())
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-50s n=%4d  %9.1f us  %5.1f%%  avg %6.1f us" % (n[-50:], a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))
print("total %.1f us" % tot)
