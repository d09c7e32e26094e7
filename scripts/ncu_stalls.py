#!/usr/bin/env python
"""Top stall locations from `ncu -i X.ncu-rep --page source --csv` (first kernel in the file)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
tot = sum(int(r[idx['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[idx[h]]) for r in data) for h in stalls}
print('by reason:', sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
order = {id(r): i for i, r in enumerate(data)}
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:n_top]:
    s = int(r[idx['# Samples']])
    st = sorted(((int(r[idx[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print('%5d %5.1f%%  #%-5d %-72s %s' % (s, 100.0 * s / tot, order[id(r)], r[idx['Source']].strip()[:72], st))
