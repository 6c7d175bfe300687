"""MD frame loop -- the caller of the hot path (util/slv_md-run_nopp:116-209, util/slv_md-run:58-169).

``run`` reads the $Frag input, streams a trajectory through ``EFP.eval_frames`` in batches and writes the report in
the reference's column format: ``# Frame EL-CAMM EL-MEA EL-EA CORR-MEA CORR-EA POL-MEA POL-EA REP-MEA REP-EA DISP-MEA
DISP-EA DISP-ISO TOTAL RMS-C RMS-S RMS-AVE`` (13 shifts %10.2f, 3 rms %10.5f).  Trajectories: a NumPy array
[nframes, natoms, 3], a ``.npy`` file, or a multi-frame ``.xyz`` file; coordinates in Angstrom are converted to Bohr
like the reference does (slv_md-run_nopp:205).  MDAnalysis readers (XTC/DCD/mdcrd) are not available here; anything
that yields per-frame coordinate arrays can be passed as an iterable.  Frames may be sharded over ranks
(``rank``/``world``): each rank evaluates its contiguous block and writes its own rows; frame numbers stay global.
"""
import sys

import numpy

from . import _lib
from .md import MDInput
from .shard import frame_range
from .solefp import EFP, REPORT_KEYS
from .units import AngstromToBohr

HEADER = '%8s %10s %10s %10s %10s %10s %10s %10s %10s %10s %10s %10s %10s %10s %10s %10s %10s' % (
    '# Frame'.ljust(8), 'EL-CAMM'.rjust(10), 'EL-MEA'.rjust(10), 'EL-EA'.rjust(10), 'CORR-MEA'.rjust(10),
    'CORR-EA'.rjust(10), 'POL-MEA'.rjust(10), 'POL-EA'.rjust(10), 'REP-MEA'.rjust(10), 'REP-EA'.rjust(10),
    'DISP-MEA'.rjust(10), 'DISP-EA'.rjust(10), 'DISP-ISO'.rjust(10), 'TOTAL'.rjust(10), 'RMS-C'.rjust(10),
    'RMS-S'.rjust(10), 'RMS-AVE'.rjust(10))


def read_xyz_frames(path):
    """Multi-frame XYZ -> float64 array [nframes, natoms, 3] (units as in the file)."""
    frames = []
    with open(path) as fh:
        while True:
            head = fh.readline()
            if not head.strip():
                break
            n = int(head.split()[0]); fh.readline()
            frames.append([[float(x) for x in fh.readline().split()[1:4]] for _ in range(n)])
    return numpy.array(frames, numpy.float64)


def format_rows(rows, shifts, first_frame=1):
    """Report lines for device output rows (slv_md-run_nopp:148-153)."""
    O = _lib.OUT
    lines = []
    for f in range(len(rows)):
        vals = tuple(float(shifts[k][f]) for k in REPORT_KEYS)
        log = "%8d" % (first_frame + f)
        log += len(vals) * " %10.2f" % vals
        log += " %10.5f %10.5f %10.5f" % (rows[f, O['rms_c']], rows[f, O['rms_smax']], rows[f, O['rms_ave']])
        lines.append(log + '\n')
    return lines


def run(solinp, traj, out, mode, nframes=None, cuts=(40.0, 17.0, 12.0), no_elect=False, no_polar=False, no_repul=False,
        no_disp=False, no_correc=False, no_ea=False, units='angstrom', batch=4096, rank=0, world=1):
    """Evaluate the SolEFP shifts of every frame and write the report; returns (rows, status) of this rank."""
    args, idx = MDInput(solinp).get()
    efp = EFP(elect=not no_elect, pol=not no_polar, rep=not no_repul, disp=not no_disp, corr=not no_correc,
              ccut=cuts[0], pcut=cuts[1], ecut=cuts[2], freq=True, cunit=True, mode=mode, ea=not no_ea)
    if isinstance(traj, str):
        traj = numpy.load(traj) if traj.endswith('.npy') else read_xyz_frames(traj)
    traj = numpy.asarray(traj, numpy.float64)
    if nframes is not None:
        traj = traj[:nframes]
    lo, hi = frame_range(len(traj), rank, world)
    scale = AngstromToBohr if units.lower().startswith('ang') else 1.0
    all_rows, all_status = [], []
    fh = open(out, 'w') if isinstance(out, str) else out
    if rank == 0:
        fh.write(HEADER + '\n'); fh.flush()
    first = True
    for b0 in range(lo, hi, batch):
        b1 = min(hi, b0 + batch)
        coords = numpy.ascontiguousarray(traj[b0:b1][:, idx, :] * scale)
        rows, status = efp.eval_frames(coords, *args) if first else efp.eval_frames(coords)
        first = False
        fh.writelines(format_rows(rows, efp.shifts_of(rows), first_frame=b0 + 1)); fh.flush()
        all_rows.append(rows.copy()); all_status.append(status.copy())
    if isinstance(out, str):
        fh.close()
    if not all_rows:
        return numpy.zeros((0, _lib.NOUT)), numpy.zeros(0, numpy.int32)
    return numpy.concatenate(all_rows), numpy.concatenate(all_status)


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description='SolEFP frequency shifts over an MD trajectory (slv_md-run on the B200)')
    ap.add_argument('-i', '--inp', required=True, help='$Frag input file')
    ap.add_argument('-t', '--traj', required=True, help='.npy [nframes,natoms,3] or multi-frame .xyz')
    ap.add_argument('-o', '--out', default='shifts.dat')
    ap.add_argument('-m', '--mode', type=int, required=True)
    ap.add_argument('-n', '--nframes', type=int, default=None)
    ap.add_argument('-c', '--ccut', type=float, default=40.0)
    ap.add_argument('-p', '--pcut', type=float, default=17.0)
    ap.add_argument('-e', '--ecut', type=float, default=12.0)
    ap.add_argument('--bohr', action='store_true', help='trajectory already in Bohr')
    for f in ('no-elect', 'no-polar', 'no-repul', 'no-disp', 'no-correc', 'no-ea'):
        ap.add_argument('--' + f, action='store_true')
    a = ap.parse_args(argv)
    run(a.inp, a.traj, a.out, a.mode, a.nframes, (a.ccut, a.pcut, a.ecut), a.no_elect, a.no_polar, a.no_repul, a.no_disp,
        a.no_correc, a.no_ea, units='bohr' if a.bohr else 'angstrom')


if __name__ == '__main__':
    main(sys.argv[1:])
