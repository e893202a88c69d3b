"""Warp-stall samples of one kernel by SASS region (200 instructions each), from `ncu -i X.ncu-rep --page source
--csv`.  Regions are labelled by their dominant opcodes, which is enough to tell the multiplier (IMAD.WIDE), the
inversion (IADD3.X / SHF / SEL), the slot operations and the kernel body apart.  python tools/ncu_regions.py src.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S, EX = ix['# Samples'], ix['Instructions Executed']
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[S]) for r in data)
print(rows[0][1])
print("total samples %d over %d SASS instructions" % (tot, len(data)))
allst = collections.Counter()
for r in data:
    for h in stalls: allst[h] += int(r[ix[h]])
print("all: " + ", ".join("%s %.1f%%" % (k.replace('stall_', ''), 100.0 * v / tot) for k, v in allst.most_common(8)))
B = 200
for b0 in range(0, len(data), B):
    chunk = data[b0:b0 + B]
    smp = sum(int(r[S]) for r in chunk)
    if smp == 0: continue
    ex = sum(int(r[EX]) for r in chunk)
    ops = collections.Counter((r[1].split()[1] if r[1].strip().startswith('@') else r[1].split()[0]) for r in chunk)
    st = collections.Counter()
    for r in chunk:
        for h in stalls: st[h] += int(r[ix[h]])
    top = ', '.join('%s %d' % (k.replace('stall_', ''), v) for k, v in st.most_common(4))
    print("%5d-%5d  %5.1f%% of samples  exec %11d  [%s]  %s" % (b0, b0 + B, 100.0 * smp / tot, ex, ' '.join('%s:%d' % kv for kv in ops.most_common(3)), top))
