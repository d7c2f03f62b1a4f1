"""Aggregate an `ncu --page source --csv` dump by SASS opcode and list the hottest instructions."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si, ei, smp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ops = collections.Counter(); tot = 0; samp = collections.Counter(); ts = 0
body = []
for r in rows[2:]:
    try: n = int(r[ei]); s = int(r[smp])
    except Exception: continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si])
    op = m.group(2).split('.')[0] if m else '?'
    ops[op] += n; tot += n; samp[op] += s; ts += s
    body.append((n, s, r[si]))
print("opcode        executed      %   stall-samples %")
for k, v in ops.most_common(28):
    print(f"{k:12s} {v/1e6:10.1f}M {100*v/tot:5.1f}%   {100*samp[k]/max(ts,1):5.1f}%")
print('total warp instructions', tot/1e9, 'G')
if len(sys.argv) > 2:
    print("hottest instructions by samples")
    for n, s, src in sorted(body, key=lambda t: -t[1])[:int(sys.argv[2])]:
        print(f"{s:7d} {n/1e6:9.1f}M  {src[:110]}")
