#!/usr/bin/env python3
"""Dev aid: per-source-line histogram of FP opcodes in an nvdisasm -gi listing range (to compare ptxas fusion decisions)."""
import re, sys
from collections import defaultdict, Counter
path, a, b, pat = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
cur = None
fresh = True
hist = defaultdict(Counter)
order = []
for l in open(path).read().split('\n')[a-1:b]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh: cur = (m.group(1).split('/')[-1], int(m.group(2))); fresh = False      # innermost frame comes first
        continue
    m = re.search(r'/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_\.]+)', l)
    if m: fresh = True
    if m and cur and pat in cur[0]:
        op = m.group(1)
        if op.split('.')[0] in ('FFMA', 'FMUL', 'FADD', 'DFMA', 'DMUL', 'DADD', 'MUFU', 'FMNMX', 'F2F', 'FSETP', 'DSETP', 'DMNMX', 'FSEL', 'TEX', 'FCHK'):
            if cur not in hist: order.append(cur)
            hist[cur][op.split('.')[0]] += 1
for k in sorted(hist, key=lambda k: k[1]):
    print(k[1], dict(hist[k]))
