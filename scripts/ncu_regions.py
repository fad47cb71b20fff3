#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump by address ranges: samples, executed instructions and stall reasons.
    python scripts/ncu_regions.py src.csv 0x0:0x1c70:prologue 0x1c70:0x5d50:acs ..."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
def addr(r):
    a = r[ix['Address']]
    return int(a, 16) if a.startswith('0x') else int(a)
base = addr(data[0])
tot = sum(f(r, '# Samples') for r in data)
print('kernel:', rows[0][1][:100]); print('total samples %.0f, warp instructions %.4g' % (tot, sum(f(r, 'Instructions Executed') for r in data)))
keys = ['stall_selected','stall_wait','stall_no_inst','stall_short_sb','stall_long_sb','stall_math','stall_not_selected','stall_dispatch','stall_branch_resolving','stall_mio','stall_lg','stall_barrier']
for spec in sys.argv[2:]:
    lo, hi, name = spec.split(':'); lo = int(lo, 16); hi = int(hi, 16)
    sel = [r for r in data if lo <= addr(r) - base < hi]
    sm = sum(f(r, '# Samples') for r in sel)
    print('%-12s samples %6.0f (%4.1f%%) instr %.4g  | ' % (name, sm, 100 * sm / tot, sum(f(r, 'Instructions Executed') for r in sel)) +
          ' '.join('%s %.1f%%' % (k[6:], 100 * sum(f(r, k) for r in sel) / max(sm, 1)) for k in keys))
if '--top' in sys.argv:
    pass
