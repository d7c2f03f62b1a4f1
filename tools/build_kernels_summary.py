"""Per-kernel time and DRAM / L2 bytes of ONE tree build from an ncu csv taken with
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv \
      --log-file out.csv python tools/prof_build.py
(tools/prof_build.py runs three builds; the last one is summarised). usage: python tools/build_kernels_summary.py out.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i
        break
ki, mi, vi, ui, idi = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
per = collections.OrderedDict()
for r in rows[start + 1:]:
    key = (r[idi], r[ki].split('(')[0].replace('void ', '').replace('rk::<unnamed>::', '').replace('rk::', '')[-44:])
    v, u = float(r[vi].replace(',', '')), r[ui]
    if r[mi] == 'gpu__time_duration.sum':
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)  # us
    else:
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    per.setdefault(key, {})[r[mi]] = v
keys = list(per.keys())
n = len(keys) // 3
agg = collections.OrderedDict()
for k in keys[2 * n:]:
    d, a = per[k], agg.setdefault(k[1], [0, 0, 0, 0, 0])
    a[0] += d['gpu__time_duration.sum']
    a[1] += d['dram__bytes_read.sum']
    a[2] += d['dram__bytes_write.sum']
    a[3] += d['lts__t_bytes.sum']
    a[4] += 1
print(f"{'kernel':44s}  n      us  dramR MB  dramW MB    L2 MB  dram GB/s   L2 GB/s")
for k, a in agg.items():
    print(f"{k:44s} {a[4]:2d} {a[0]:7.1f} {a[1]/1e6:9.1f} {a[2]/1e6:9.1f} {a[3]/1e6:8.1f} {(a[1]+a[2])/a[0]/1e3:10.0f} {a[3]/a[0]/1e3:9.0f}")
print('total us (cold cache, serialised under ncu)', round(sum(a[0] for a in agg.values()), 1))
