#!/usr/bin/env python
"""Collect the reference's own known-answer data for the SolEFP path into tests/golden/small_cluster.json:
  examples/1_small-clusters/nma.xyz, water.xyz  (input geometries, Angstrom)
  examples/1_small-clusters/f                   (full-precision shift dictionary of scan point 0)
  examples/1_small-clusters/dat                 (21-point scan: total, ele_tot at 2 decimals)
Run in the build container (reads /root/reference); the JSON is the committed fixture."""
import ast, json, os, sys
src = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/examples/1_small-clusters'
def xyz(fn):
    rows = []
    for ln in open(os.path.join(src, fn)).read().split('\n')[2:]:
        p = ln.split()
        if len(p) >= 4:
            rows.append([p[0]] + [float(x) for x in p[1:4]])
    return rows
ftxt = open(os.path.join(src, 'f')).read()
gold = ast.literal_eval(ftxt[ftxt.index('{'):ftxt.index('}') + 1])
dat = [[float(x) for x in ln.split()] for ln in open(os.path.join(src, 'dat')) if ln.strip()]
out = dict(source='globulion/slv examples/1_small-clusters (scan.py recipe: mode 8, no cut-offs, supl [2,8,3,5]/[1,2,3], '
                  'water 6 Bohr along N->H, 20 rotations of 0.1 rad about z)',
           nma_xyz=xyz('nma.xyz'), water_xyz=xyz('water.xyz'), f=gold, dat=dat)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'small_cluster.json'), 'w'), indent=1)
print('ok', len(dat), sorted(gold)[:3])
