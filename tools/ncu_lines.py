"""Per-CUDA-source-line executed warp instructions from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
per = collections.OrderedDict(); cur = None; tot = 0; ts = 0
for r in rows:
    if len(r) > 8 and r[0] == 'Line No':
        hdr = r; ei = r.index('Instructions Executed'); smp = r.index('# Samples'); continue
    if hdr is None or len(r) <= ei: continue
    if r[0].strip():
        cur = (r[0], r[1].strip())
        per.setdefault(cur, [0, 0])
    elif cur is not None:
        try: n = int(r[ei]); s = int(r[smp])
        except ValueError: continue
        per[cur][0] += n; per[cur][1] += s; tot += n; ts += s
print('total', tot / 1e9, 'G warp instructions')
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for (ln, src), (n, s) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:>5s} {n/1e6:9.1f}M {100*n/tot:5.1f}%  smp {100*s/max(ts,1):5.1f}%  {src[:105]}")
