"""Aggregate an ncu source page (ncu -i rep --page source --csv --print-source cuda,sass -k ... ) by source line."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.Counter(); samp = collections.Counter(); thr = collections.Counter(); lines = {}
cur = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; hdr = None; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No':
        hdr = r; li = 0; ad = hdr.index('Address'); ie = hdr.index('Instructions Executed'); ws = hdr.index('Warp Stall Sampling (All Samples)'); te = hdr.index('Thread Instructions Executed'); continue
    if hdr is None or cur is None or r[ad] != '-' or not r[li]: continue
    ln = int(r[li]); agg[(cur, ln)] += float(r[ie] or 0); samp[(cur, ln)] += float(r[ws] or 0); thr[(cur, ln)] += float(r[te] or 0); lines[(cur, ln)] = r[1].strip()[:100]
tot = sum(agg.values()); ts = sum(samp.values())
print('total warp-instr', tot, 'samples', ts)
byfile = collections.Counter(); sf = collections.Counter()
for (f, l), n in agg.items(): byfile[f] += n; sf[f] += samp[(f, l)]
print({f: (round(n / tot, 3), round(sf[f] / ts, 3)) for f, n in byfile.most_common()})
for (f, l), n in agg.most_common(top):
    print(f"{f}:{l:4d} inst {n/tot*100:5.2f}% samp {samp[(f,l)]/ts*100:5.2f}% thr/inst {thr[(f,l)]/max(n,1):5.1f}  {lines[(f,l)]}")
