#!/usr/bin/env python3
"""Print a PTX range with .loc lines folded into short `// file:line` comments (dev aid)."""
import re, sys
path, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
names = {}
lines = open(path).read().split('\n')
for l in lines:
    m = re.match(r'\s*\.file\s+(\d+)\s+"([^"]+)"', l)
    if m: names[m.group(1)] = m.group(2).split('/')[-1]
cur = ''
for i in range(a-1, min(b, len(lines))):
    l = lines[i]
    m = re.match(r'\s*\.loc\s+(\d+)\s+(\d+)\s+(\d+)(.*)', l)
    if m:
        inl = re.search(r'inlined_at\s+(\d+)\s+(\d+)', m.group(4))
        cur = f"{names.get(m.group(1), m.group(1))}:{m.group(2)}" + (f" <-{inl.group(2)}" if inl else '')
        continue
    if l.strip() == '' : continue
    print(f"{i+1:6d} {l.strip():70s} // {cur}")
    cur = ''
