#!/usr/bin/env python
"""The reference's small-cluster example on the B200 path: NMA amide-I shift against one water molecule placed
6 Bohr along N->H and rotated about z in 20 steps of 0.1 rad (same recipe and same API calls as
examples/1_small-clusters/scan.py of globulion/slv; libbbg's QMFile is replaced by a five-line xyz reader).
Writes `dat` (total, ele_tot per scan point).  Needs a GPU."""
import os
import sys

import numpy
from numpy import cos, sin

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from solvshift.slvpar import Frag        # noqa: E402
from solvshift.solefp import EFP         # noqa: E402

BohrToAngstrom = 0.529177249
here = os.path.dirname(os.path.abspath(__file__))


def read_xyz(fn):
    rows = [ln.split() for ln in open(os.path.join(here, fn)).read().split('\n')[2:] if len(ln.split()) >= 4]
    return numpy.array([[float(x) for x in r[1:4]] for r in rows]) / BohrToAngstrom


xyz_nma = read_xyz('nma.xyz'); xyz_wat = read_xyz('water.xyz')
xyz_wat -= xyz_wat.mean(axis=0); xyz_nma -= xyz_nma.mean(axis=0)
r = xyz_nma[7] - xyz_nma[1]
xyz_wat += 6.0 * r / numpy.linalg.norm(r)

mode = 8
supl = [numpy.array([2, 8, 3, 5]) - 1, numpy.array([1, 2, 3]) - 1]
reord = [numpy.arange(12), numpy.arange(3)]
frg_solute, frg_solvent = Frag('nma'), Frag('water')
nmol = [frg_solute.get_natoms(), frg_solvent.get_natoms()]
ind = [0, 1]
bsm = (frg_solute, frg_solvent)

efp = EFP(elect=1, pol=1, disp=1, rep=1, corr=1, freq=True, mode=mode, ccut=10e10, pcut=10e10, ecut=10e10, cunit=True)

t = 0.1
rot_mat = numpy.array([cos(t), -sin(t), 0, sin(t), cos(t), 0, 0, 0, 1]).reshape(3, 3)
results = []
for i in range(21):
    xyz = numpy.concatenate([xyz_nma, xyz_wat])
    efp.set(xyz, ind, nmol, bsm, supl, reord)
    efp.eval(0)
    shifts = efp.get_shifts()
    results.append([shifts['total'], shifts['ele_tot']])
    if i == 0:
        print(shifts)
    xyz_wat = numpy.dot(xyz_wat, rot_mat)

with open(os.path.join(here, 'dat'), 'w') as out:
    for tot, ele in results:
        out.write('%14.2f %14.2f\n' % (tot, ele))
print(numpy.array(results))
print(efp.get_rms_max())
