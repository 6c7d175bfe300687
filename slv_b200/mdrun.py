"""MD frame loop -- the caller of the hot path (util/slv_md-run_nopp:116-209, util/slv_md-run:58-169).

``run`` reads the $Frag input, streams a trajectory through ``EFP.eval_frames`` in batches and writes the report in
the reference's column format: ``# Frame EL-CAMM EL-MEA EL-EA CORR-MEA CORR-EA POL-MEA POL-EA REP-MEA REP-EA DISP-MEA
DISP-EA DISP-ISO TOTAL RMS-C RMS-S RMS-AVE`` (13 shifts %10.2f, 3 rms %10.5f).  Trajectories: a NumPy array
[nframes, natoms, 3], a ``.npy`` file, or a multi-frame ``.xyz`` file; coordinates in Angstrom are converted to Bohr
like the reference does (slv_md-run_nopp:205).  DCD (NAMD/CHARMM) and AMBER mdcrd
files and GROMACS XTC (xdr3dcoord-compressed, nm) are read natively (the reference goes through MDAnalysis / its own
readers, util/slv_md-run:117-152); float32 records go to the device as they are (gather + unit conversion there).  Frames may be sharded over ranks
(``rank``/``world``): each rank evaluates its contiguous block and writes its own rows; frame numbers stay global.
"""
import sys
import warnings

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


def read_dcd_frames(path, first=0, last=None):
    """CHARMM/NAMD DCD trajectory (the `namd-charmm` branch of util/slv_md-run:117-152) -> float32 [nframes, natoms, 3]
    in the file's units (Angstrom), the precision the file holds.  Little- or big-endian, with or without unit-cell records, fixed atoms unsupported."""
    import struct
    with open(path, 'rb') as fh:
        raw = fh.read()
    en = '<' if struct.unpack('<i', raw[:4])[0] == 84 else '>'
    if struct.unpack(en + 'i', raw[:4])[0] != 84 or raw[4:8] != b'CORD':
        raise IOError('%s is not a DCD coordinate file' % path)
    icntrl = struct.unpack(en + '20i', raw[8:88])
    nset, has_cell, nfixed = icntrl[0], icntrl[10] != 0, icntrl[8]
    if nfixed:
        raise NotImplementedError('DCD files with fixed atoms are not supported')
    pos = 92
    n = struct.unpack(en + 'i', raw[pos:pos + 4])[0]; pos += 8 + n            # title block
    pos += 4; natoms = struct.unpack(en + 'i', raw[pos:pos + 4])[0]; pos += 8  # natoms block
    frame_bytes = (56 if has_cell else 0) + 3 * (8 + 4 * natoms)
    nset = min(nset, (len(raw) - pos) // frame_bytes) if nset > 0 else (len(raw) - pos) // frame_bytes
    last = nset if last is None else min(last, nset)
    out = numpy.empty((max(0, last - first), natoms, 3), numpy.float32)      # as stored: float32 records
    for f in range(first, last):
        q = pos + f * frame_bytes + (56 if has_cell else 0)
        for d in range(3):
            out[f - first, :, d] = numpy.frombuffer(raw, en + 'f4', natoms, q + 4)
            q += 8 + 4 * natoms
    return out


def read_mdcrd_frames(path, natoms, box=False):
    """AMBER ASCII trajectory (the `amber` branch of util/slv_md-run:128-130): title line, then 10F8.3 records;
    `box` skips the three box lengths after every frame."""
    with open(path) as fh:
        fh.readline()
        vals = numpy.array(fh.read().split(), numpy.float64)
    per = 3 * natoms + (3 if box else 0)
    nfr = len(vals) // per
    return vals[:nfr * per].reshape(nfr, per)[:, :3 * natoms].reshape(nfr, natoms, 3).copy()


def read_xtc_frames(path, first=0, last=None):
    """GROMACS XTC trajectory (the `gromacs` branch of util/slv_md-run:117-152) -> float32 [nframes, natoms, 3] in nm, the
    records the file holds; decoded by the library (slv_b200/csrc/xtc_reader.h)."""
    import ctypes
    n = ctypes.c_int64(0); na = ctypes.c_int32(0)
    _lib.check(_lib.lib.slvgpu_xtc_scan(path.encode(), ctypes.byref(n), ctypes.byref(na)))
    last = n.value if last is None else min(last, n.value)
    cnt = max(0, last - first)
    out = numpy.empty((cnt, na.value, 3), numpy.float32)
    if cnt:
        _lib.check(_lib.lib.slvgpu_xtc_read(path.encode(), first, cnt, out.ctypes.data, None, None, None))
    return out


def load_trajectory(traj, natoms=None):
    """numpy array, .npy, multi-frame .xyz, .dcd, GROMACS .xtc (nm) or AMBER .mdcrd/.crd (needs natoms)."""
    if not isinstance(traj, str):
        traj = numpy.asarray(traj)
        return traj if traj.dtype == numpy.float32 else numpy.asarray(traj, numpy.float64)
    low = traj.lower()
    if low.endswith('.npy'):
        return numpy.load(traj)
    if low.endswith('.dcd'):
        return read_dcd_frames(traj)
    if low.endswith('.xtc'):
        return read_xtc_frames(traj)
    if low.endswith('.mdcrd') or low.endswith('.crd'):
        assert natoms is not None, 'AMBER trajectories need the number of atoms'
        return read_mdcrd_frames(traj, natoms)
    return read_xyz_frames(traj)


def status_text(st):
    """Names of the per-frame status bits (include/slvgpu.h SLVGPU_ST_*)."""
    names = [(_lib.ST_NONFINITE, 'non-finite'), (_lib.ST_POL_NOCONV, 'pol-not-converged'), (_lib.ST_SINGULAR, 'singular-fit'),
             (_lib.ST_OVERFLOW, 'pol-workspace-overflow')]
    return '+'.join(n for b, n in names if st & b) or 'ok'


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
        no_disp=False, no_correc=False, no_ea=False, units='angstrom', batch=4096, rank=0, world=1, first_frame=0,
        append=False, natoms=None):
    """Evaluate the SolEFP shifts of every frame and write the report; returns (rows, status) of this rank.
    `first_frame` (0-based) + `append` resume an interrupted run: frames are independent, so a restart just skips the
    rows already in the report (the reference appends and flushes per frame but has no restart flag)."""
    args, idx = MDInput(solinp).get()
    efp = EFP(elect=not no_elect, pol=not no_polar, rep=not no_repul, disp=not no_disp, corr=not no_correc,
              ccut=cuts[0], pcut=cuts[1], ecut=cuts[2], freq=True, cunit=True, mode=mode, ea=not no_ea)
    traj_name = traj
    traj = load_trajectory(traj, natoms)
    f32 = traj.dtype == numpy.float32          # trajectory records as the MD engine wrote them: gather + units on the device
    if not f32:
        traj = numpy.asarray(traj, numpy.float64)
    if nframes is not None:
        traj = traj[:nframes]
    lo, hi = frame_range(len(traj) - first_frame, rank, world)
    lo += first_frame; hi += first_frame
    if isinstance(traj_name, str) and traj_name.lower().endswith('.xtc') and units.lower().startswith('ang'):
        units = 'nm'                           # XTC records are in nm whatever the caller assumed for the other formats
    scale = {'ang': AngstromToBohr, 'nm': 10.0 * AngstromToBohr}.get(units.lower()[:3] if units.lower().startswith('ang') else units.lower(), 1.0)
    all_rows, all_status = [], []
    if world > 1 and isinstance(out, str):
        # every rank writes its own block of frames: one file per rank (the ranks would clobber a shared path)
        out = '%s.rank%d' % (out, rank)
    fh = None if out is None else (open(out, 'a' if append else 'w') if isinstance(out, str) else out)
    if fh is not None and (rank == 0 or world > 1) and not append:
        fh.write(HEADER + '\n'); fh.flush()
    first = True
    for b0 in range(lo, hi, batch):
        b1 = min(hi, b0 + batch)
        if f32:
            kw = dict(scale=scale, idx=numpy.asarray(idx, numpy.int32))
            coords = numpy.ascontiguousarray(traj[b0:b1])
        else:
            kw = {}
            coords = numpy.ascontiguousarray(traj[b0:b1][:, idx, :] * scale)
        rows, status = efp.eval_frames(coords, *args, **kw) if first else efp.eval_frames(coords, **kw)
        first = False
        bad = numpy.nonzero(numpy.asarray(status) != 0)[0]
        if len(bad):
            # a flagged frame is not an ordinary number: non-finite term, polarisation solve not converged, or more
            # centres inside pcut than the workspace holds (pol_mea = 0 then); say so next to the rows
            msg = ', '.join('%d:%s' % (b0 + 1 + int(f), status_text(int(status[f]))) for f in bad[:8])
            warnings.warn('%d frame(s) of this batch carry a non-zero status (%s%s)' % (len(bad), msg, ', ...' if len(bad) > 8 else ''))
        if fh is not None:
            lines = format_rows(rows, efp.shifts_of(rows), first_frame=b0 + 1)
            for f in bad:
                lines[f] = lines[f].rstrip('\n') + '   # status: %s\n' % status_text(int(status[f]))
            fh.writelines(lines); fh.flush()
        all_rows.append(rows.copy()); all_status.append(status.copy())
    if isinstance(out, str):
        fh.close()
    run.last_efp = efp
    if not all_rows:
        return numpy.zeros((0, _lib.NOUT)), numpy.zeros(0, numpy.int32)
    return numpy.concatenate(all_rows), numpy.concatenate(all_status)


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description='SolEFP frequency shifts over an MD trajectory (slv_md-run on the B200)')
    ap.add_argument('-i', '--inp', required=True, help='$Frag input file')
    ap.add_argument('-t', '--traj', required=True, help='.npy [nframes,natoms,3], multi-frame .xyz, .dcd, .xtc or AMBER .mdcrd')
    ap.add_argument('--natoms', type=int, default=None, help='atoms per frame (AMBER trajectories)')
    ap.add_argument('--first-frame', type=int, default=0, help='0-based frame to start from (resume)')
    ap.add_argument('--append', action='store_true', help='append to an existing report (resume)')
    ap.add_argument('-o', '--out', default='shifts.dat')
    ap.add_argument('-m', '--mode', type=int, required=True)
    ap.add_argument('-n', '--nframes', type=int, default=None)
    ap.add_argument('-c', '--ccut', type=float, default=40.0)
    ap.add_argument('-p', '--pcut', type=float, default=17.0)
    ap.add_argument('-e', '--ecut', type=float, default=12.0)
    ap.add_argument('--bohr', action='store_true', help='trajectory already in Bohr')
    for f in ('no-elect', 'no-polar', 'no-repul', 'no-disp', 'no-correc', 'no-ea'):
        ap.add_argument('--' + f, action='store_true')
    ap.add_argument('--distributed', action='store_true',
                    help='under torchrun: frames are sharded over the ranks (one GPU each), rank 0 gathers the rows and writes the report')
    a = ap.parse_args(argv)
    if a.distributed:
        return _main_distributed(a)
    run(a.inp, a.traj, a.out, a.mode, a.nframes, (a.ccut, a.pcut, a.ecut), a.no_elect, a.no_polar, a.no_repul, a.no_disp,
        a.no_correc, a.no_ea, units='bohr' if a.bohr else 'angstrom', first_frame=a.first_frame, append=a.append,
        natoms=a.natoms)


def _main_distributed(a):
    import os
    import torch
    import torch.distributed as dist
    from .shard import gather_rows
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    dist.init_process_group('nccl' if cuda else 'gloo')
    traj = load_trajectory(a.traj, a.natoms)
    if a.nframes is not None:
        traj = traj[:a.nframes]
    rows, status = run(a.inp, traj, None, a.mode, None, (a.ccut, a.pcut, a.ecut), a.no_elect, a.no_polar, a.no_repul, a.no_disp,
                       a.no_correc, a.no_ea, units='bohr' if a.bohr else 'angstrom', rank=rank, world=world)
    t = torch.from_numpy(rows)
    if cuda:
        t = t.cuda()
    allrows = gather_rows(t, len(traj), dst=0)          # the run's one collective
    if rank == 0:
        allrows = allrows.cpu().numpy()
        with open(a.out, 'w') as fh:
            fh.write(HEADER + '\n')
            fh.writelines(format_rows(allrows, run.last_efp.shifts_of(allrows), first_frame=1))
    dist.destroy_process_group()


if __name__ == '__main__':
    main(sys.argv[1:])
