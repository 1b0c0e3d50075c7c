#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` listing per CUDA source line (samples / instructions)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; iS = hdr.index('# Samples'); iE = hdr.index('Instructions Executed'); continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    if not r[0].strip().isdigit(): continue
    # a CUDA line row: Line No, Source, then aggregated metrics
    try: s = int(r[iS] or 0); e = int(r[iE] or 0)
    except ValueError: continue
    if r[2].startswith('0x'): continue          # SASS child rows
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1].strip()]); a[0] += s; a[1] += e
ts = sum(a[0] for a in agg.values()) or 1; te = sum(a[1] for a in agg.values()) or 1
print('total samples', ts, 'inst', te)
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%-18s %5d  samp %5.1f%%  inst %5.1f%%  %s' % (f, ln, 100 * a[0] / ts, 100 * a[1] / te, a[2][:110]))
