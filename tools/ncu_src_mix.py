"""Instruction mix / stall-sample summary of one kernel from `ncu --page source --csv`."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
out = []
blocks = []   # one block per kernel instance
cur = None
for r in rows:
    if r and r[0] == 'Address':
        cur = dict(h=r, data=[])
        blocks.append(cur)
    elif cur is not None and r and r[0].startswith('0x'):
        cur['data'].append(r)
b = blocks[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
h, data = b['h'], b['data']
si, samp, ex = h.index('Source'), h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed')
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_')]
tot = sum(float(r[samp] or 0) for r in data)
ops, opsamp = collections.Counter(), collections.Counter()
for r in data:
    toks = r[si].split()
    op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '')
    op = op.split('.')[0]
    ops[op] += float(r[ex] or 0)
    opsamp[op] += float(r[samp] or 0)
ti = sum(ops.values())
print(f"instances {len(blocks)}; total warp-inst {ti:.0f}; samples {tot:.0f}")
for op, c in ops.most_common(22):
    print(f"{op:10s} inst {100*c/ti:5.1f}%  samples {100*opsamp[op]/max(tot,1):5.1f}%")
if stall_cols:
    st = collections.Counter()
    for r in data:
        for i in stall_cols:
            st[h[i]] += float(r[i] or 0)
    s = sum(st.values())
    print('stalls:', ', '.join(f"{k[6:]} {100*v/s:.0f}%" for k, v in st.most_common(8)))
top = sorted(data, key=lambda r: -float(r[samp] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 12]
for r in top:
    print(f"{100*float(r[samp] or 0)/max(tot,1):5.1f}%  {r[si].strip()[:90]}")
