"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr = r; start = i; break
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict(); cnt = collections.Counter()
for r in rows[start + 1:]:
    name = r[ki].split('(')[0][-52:]; v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    agg[name] = agg.get(name, 0) + v; cnt[name] += 1
tot = sum(agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:30]:
    print(f"{v/div:10.1f} us {100*v/tot:5.1f}%  n={cnt[k]/div:6.1f}  {k}")
print('total us', tot / div, 'launches', sum(cnt.values()) / div)
