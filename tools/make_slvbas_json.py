#!/usr/bin/env python
"""Convert the GAMESS-format basis table of the reference (dat/slvbas: 6-311++G** for
H Li C N O F Na S Cl) into slv_b200/data/slvbas.json: {symbol: [[type, [[exp, c1(, c2)], ...]], ...]}.
Shell types S, P, D, L (=SP).  Run once in the build container; the JSON is the committed fixture."""
import json, re, sys
src = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/dat/slvbas'
out = sys.argv[2] if len(sys.argv) > 2 else 'slv_b200/data/slvbas.json'
basis = {}; cur = None; lines = open(src).read().split('\n'); i = 0
while i < len(lines):
    ln = lines[i].split()
    if len(ln) == 2 and ln[1] == 'slvbas':
        cur = ln[0]; basis[cur] = []; i += 1; continue
    if len(ln) == 2 and ln[0] in ('S', 'P', 'D', 'L') and ln[1].isdigit():
        n = int(ln[1]); prims = []
        for k in range(n):
            p = lines[i + 1 + k].split()
            prims.append([float(x) for x in p[1:]])
        basis[cur].append([ln[0], prims]); i += n + 1; continue
    i += 1
json.dump(basis, open(out, 'w'), indent=0, separators=(',', ':'))
print({k: [s[0] + str(len(s[1])) for s in v] for k, v in basis.items()})
