#!/usr/bin/env python3
"""Dev aid: print the arithmetic op sequence (opcode + immediate operands) of a PTX line range, registers abstracted,
so two compilations of the same formula can be diffed."""
import re, sys
path, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
skip = ('mov.', 'ld.param', 'cvta', 'st.local', 'ld.local', 'bra', '.loc', '@', 'setp', 'selp', 'and.', 'or.', 'shl', 'shr', 'xor', 'ld.global', 'st.global', 'ld.const')
for i, l in enumerate(open(path).read().split('\n')[a-1:b]):
    l = l.strip()
    if not l or l.startswith('//') or l.startswith('$') or l.startswith('{') or l.startswith('}') or l.startswith('.'): continue
    if l.startswith(skip): continue
    m = re.match(r'([\w\.]+)\s+(.*);', l)
    if not m: continue
    op, args = m.group(1), m.group(2)
    imm = [x for x in re.findall(r'0[fd][0-9A-Fa-f]+', args)]
    print(op, ' '.join(imm))
