"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the LAST MSM
(from its k_from_mont launch on) and the sequence of the big launches.  python tools/launch_summary.py file.csv"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
rows = rows[hdr + 1:]
names = [re.sub(r'\(.*', '', r[4]).replace('void ', '').replace('mnt753::', '') for r in rows]
# an MSM starts with its (possibly chunked) k_from_mont / k_count launches and has exactly one k_scatter
end = max(i for i, n in enumerate(names) if n.startswith('k_horner'))     # last COMPLETE MSM
last_scatter = max(i for i, n in enumerate(names[:end]) if n.startswith('k_scatter'))
s = last_scatter
while s > 0 and (names[s - 1].startswith('k_from_mont') or names[s - 1].startswith('k_count') or names[s - 1].startswith('k_scan')):
    s -= 1
agg, seq, tot = {}, [], 0.0
for r, n in zip(rows[s:end + 1], names[s:end + 1]):
    t = float(r[-1]) / 1e6
    tot += t
    agg[n] = agg.get(n, 0) + t
    seq.append((n, t, r[8], r[7]))
print("last MSM: %d launches, %.3f ms summed" % (len(seq), tot))
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    print("%9.3f ms  %5.1f%%  %s" % (v, 100 * v / tot, k))
print()
for n, t, g, b in seq:
    if t > 0.05:
        print("%9.3f ms  grid %-14s block %-12s %s" % (t, g, b, n))
