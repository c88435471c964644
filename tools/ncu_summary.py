#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` export: instruction totals, stall
reasons and the most-sampled instructions.  usage: ncu_summary.py sass.csv [items]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
items = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
hdr = rows[1]
ie, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
secs = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
end = secs[1] if len(secs) > 1 else len(rows)
body = rows[2:end]
tot = sum(int(r[ie]) for r in body)
print(rows[0][1][:80])
print('total warp-instructions', tot, 'per item', round(tot / items, 1), 'sass lines', len(body))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
acc = {}
for r in body:
    for i in stall_cols:
        acc[hdr[i]] = acc.get(hdr[i], 0) + int(r[i] or 0)
s = sum(acc.values())
print('stall samples:', ', '.join(f'{k[6:]} {v / s:.1%}' for k, v in sorted(acc.items(), key=lambda kv: -kv[1])[:9]))
print('most sampled instructions:')
for r in sorted(body, key=lambda r: -int(r[isamp] or 0))[:14]:
    st = {hdr[i][6:]: r[i] for i in stall_cols if int(r[i] or 0) > 100}
    print(f'  {r[1].strip()[:64]:64s} exec={r[ie]:>9s} samples={r[isamp]:>6s} {st}')
