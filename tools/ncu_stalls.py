"""Stall-reason samples per code region from `ncu --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_stalls.py src.csv "name:lo-hi" ...   (CUDA source line ranges; lines outside = 'other')"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
regions = []
for a in sys.argv[2:]:
    n, r = a.split(':'); lo, hi = r.split('-'); regions.append((n, int(lo), int(hi)))
hdr = None; cur = None
acc = collections.defaultdict(lambda: collections.Counter())
for r in rows:
    if len(r) > 8 and r[0] == 'Line No':
        hdr = r; cols = {n: i for i, n in enumerate(r) if n.startswith('stall_') and 'Not Issued' not in n}; ei = r.index('Instructions Executed'); continue
    if hdr is None: continue
    if r[0].strip():
        try: cur = int(r[0])
        except ValueError: cur = None
        continue
    if cur is None: continue
    reg = 'other'
    for n, lo, hi in regions:
        if lo <= cur <= hi: reg = n; break
    for n, i in cols.items():
        try: acc[reg][n] += int(r[i])
        except (ValueError, IndexError): pass
    try: acc[reg]['_inst'] += int(r[ei])
    except ValueError: pass
tot = sum(sum(v for k, v in c.items() if k != '_inst') for c in acc.values())
names = sorted({k for c in acc.values() for k in c if k != '_inst'}, key=lambda k: -sum(c[k] for c in acc.values()))[:9]
print(f"{'region':22s} {'inst%':>6s} {'smp%':>6s}  " + ' '.join(f"{n[6:14]:>8s}" for n in names))
ti = sum(c['_inst'] for c in acc.values())
for reg, c in sorted(acc.items(), key=lambda kv: -sum(v for k, v in kv[1].items() if k != '_inst')):
    s = sum(v for k, v in c.items() if k != '_inst')
    print(f"{reg:22s} {100*c['_inst']/ti:6.1f} {100*s/tot:6.1f}  " + ' '.join(f"{100*c[n]/tot:8.1f}" for n in names))
