"""SolEFP for biomolecules: the protein fragmentation front-end of the frame loop
(``solvshift.biomol.BiomoleculeFragmentation``, solvshift/biomol.py:21-546; example examples/run_biomol.py).

What the reference does per trajectory (biomol.py:150-202): detect the backbone amide units, write a ``$Frag`` input
from the residue -> EFP-fragment tables of a topology-conversion (``.tc``) file (dat/gmx.tc, dat/probe.tc), parse it
with ``MDInput``; then per frame: SolCAMM of the probe against the 2 x 4 point charges of the near-zone CONH groups
(``EFP.eval_dma``), the SolEFP shifts of the far zone with ``remove_clashes=True``, and one report line.

Here the first part is host-side bookkeeping done once (pure index work, no arithmetic on coordinates beyond the
amide-detection distances), and the per-frame part is ONE batched call: the near-zone charges become one-atom
environment fragments riding on their frame atoms (``solefp.point_charge_fragment``), so ``EFP.eval_frames`` returns
the ele_env_* / cor_env_* terms next to the SolEFP terms for every frame, with the clash-removal fragment classes
resolved at plan time.  Topologies are read from GROMACS ``.gro`` or PDB files (the reference goes through MDAnalysis).
"""
import re

import numpy

from . import mdrun
from .md import MDInput
from .slvpar import Frag
from .solefp import EFP, point_charge_fragment
from .units import AngstromToBohr

__all__ = ['BiomoleculeFragmentation', 'Topology', 'parse_tc']

# ESP-fitted charges of a CONH group (C, O, N, H), from N-methylacetamide in vacuo (biomol.py:452)
CONH_CHARGES = (4.846770E-01, -5.688321E-01, -3.846247E-02, 1.192387E-01)
DW_LIST = ('solcamm', 'ele_mea', 'ele_ea', 'cor_mea', 'cor_ea', 'pol_mea', 'pol_ea', 'rep_mea', 'rep_ea', 'dis_mea', 'dis_ea',
           'dis_mea_iso', 'total')
DW_ENV_LIST = ('ele_env_mea', 'ele_env_ea', 'cor_env_mea', 'cor_env_ea', 'tot_env', 'total+env')
# backbone models of a far-zone peptide group: fragment, reorder (parameter atom <- k-th atom counted from the carbonyl C,
# 0 = absent), supl; (with N-H, next to proline) -- biomol.py:268-306
_CONH_MODELS = {
    'nma': (12, [0, 3, 1, 0, 2, 0, 0, 4, 0, 0, 0, 0], [3, 5, 2, 8], [0, 3, 1, 0, 2, 0, 0, 0, 0, 0, 0, 0], [3, 5, 2]),
    'chonh2': (6, [1, 2, 0, 3, 0, 4], [1, 2, 4, 6], [1, 2, 0, 3, 0, 0], [1, 2, 4]),
    'chonhme': (9, [1, 3, 2, 0, 0, 4, 0, 0, 0], [1, 2, 3, 6], [1, 3, 2, 0, 0, 0, 0, 0, 0], [1, 2, 3]),
    'comenh2': (9, [1, 2, 0, 3, 0, 4, 0, 0, 0], [1, 2, 4, 6], [1, 2, 0, 3, 0, 0, 0, 0, 0], [1, 2, 4]),
}


class Topology(object):
    """Per-atom residue number, residue name and atom name (what the reference takes from an MDAnalysis Universe)."""

    def __init__(self, resid, resname, name):
        self.resid = numpy.asarray(resid, int); self.resname = numpy.asarray(resname, str); self.name = numpy.asarray(name, str)
        assert len(self.resid) == len(self.resname) == len(self.name)

    def __len__(self):
        return len(self.name)

    def residues(self):
        """[(resname, resid, atom indices)] in file order (consecutive atoms with the same residue number and name)."""
        out = []; start = 0
        for i in range(1, len(self) + 1):
            if i == len(self) or self.resid[i] != self.resid[start] or self.resname[i] != self.resname[start]:
                out.append((str(self.resname[start]), int(self.resid[start]), numpy.arange(start, i)))
                start = i
        return out

    @classmethod
    def from_gro(cls, path):
        """GROMACS .gro: title, atom count, fixed columns resid(5) resname(5) atomname(5) atomnr(5) x y z [nm]."""
        with open(path) as fh:
            lines = fh.read().split('\n')
        n = int(lines[1])
        rid, rn, an, xyz = [], [], [], []
        for ln in lines[2:2 + n]:
            rid.append(int(ln[0:5])); rn.append(ln[5:10].strip()); an.append(ln[10:15].strip())
            xyz.append([float(ln[20:28]), float(ln[28:36]), float(ln[36:44])])
        return cls(rid, rn, an), numpy.array(xyz) * 10.0

    @classmethod
    def from_pdb(cls, path):
        """PDB ATOM / HETATM records; coordinates in Angstrom."""
        rid, rn, an, xyz = [], [], [], []
        with open(path) as fh:
            for ln in fh:
                if ln.startswith(('ATOM', 'HETATM')):
                    an.append(ln[12:16].strip()); rn.append(ln[17:21].strip()); rid.append(int(ln[22:26]))
                    xyz.append([float(ln[30:38]), float(ln[38:46]), float(ln[46:54])])
        return cls(rid, rn, an), numpy.array(xyz)

    @classmethod
    def read(cls, path):
        return cls.from_gro(path) if path.lower().endswith('.gro') else cls.from_pdb(path)


def parse_tc(path):
    """SolEFP-MD topology conversion file (biomol.py:478-508): ``[ RES ]`` sections holding ``@fragment`` blocks with a
    ``reorder`` and a ``supl`` line each; '#' starts a comment line.  Returns {residue: [(fragment, reorder, supl), ...]}
    (reorder / supl 1-based as in the file, 0 = atom absent from the MD residue)."""
    with open(path) as fh:
        text = '\n'.join(ln for ln in fh.read().split('\n') if not ln.strip().startswith('#'))
    res = {}
    for sec in text.split('[')[1:]:
        blocks = sec.split('@')
        name = blocks[0].split()[0]
        frags = []
        for blk in blocks[1:]:
            lines = blk.split('\n')
            fname = lines[0].strip(); reorder = supl = None
            for ln in lines[1:]:
                w = ln.split()
                if w and w[0] == 'reorder':
                    reorder = numpy.array(w[1:], int)
                elif w and w[0] == 'supl':
                    supl = numpy.array(w[1:], int)
            if reorder is None:
                raise ValueError(' No reordering list (reorder) provided for fragment < %s > in residue < %s > !' % (fname, name))
            if supl is None:
                raise ValueError(' No superimposition list (supl) provided for fragment < %s > in residue < %s > !' % (fname, name))
            frags.append((fname, reorder, supl))
        res[name] = frags
    return res


def _frag_block(res_name, frag, atom_numbers):
    name, reorder, supl = frag
    return ('$Frag\n%-10s %-10s\n%-10s %s\n%-10s %s\n%-10s %s    1\n$endFrag\n\n'
            % (res_name, name, 'reorder', ' '.join('%i' % x for x in reorder), 'supl', ','.join('%i' % x for x in supl),
               'atoms', ','.join('%i' % x for x in atom_numbers)))


class BiomoleculeFragmentation(object):
    """``method = BiomoleculeFragmentation(res, probe, mode, ...); method.run(top, traj, nframes=10, conh='nma')`` --
    constructor and ``run`` keywords as in the reference (biomol.py:109-150)."""

    def __init__(self, res, probe, mode, elect=True, polar=True, repul=True, disp=True, correc=True,
                 ccut=20.0, pcut=13.0, ecut=11.0, solcamm=None, lprint=False, is_probe_terminal=False,
                 report='report.dat', write_solefp_input=False, write_debug_file=False,
                 rcl_algorithm='remove_by_name', include_ions=True, include_polar=True):
        if solcamm is not None:
            raise NotImplementedError('an external SolCAMM DMA file is not supported: the probe fragment\'s own SolCAMM is used')
        self.report = report; self.lprint = lprint; self.write_solefp_input = write_solefp_input
        self.ccut, self.pcut, self.ecut = ccut, pcut, ecut
        self.rcl_algorithm, self.include_ions, self.include_polar = rcl_algorithm, include_ions, include_polar
        self._res = parse_tc(res) if isinstance(res, str) else dict(res)
        self._probe = parse_tc(probe) if isinstance(probe, str) else dict(probe)
        self._probe_label = list(self._probe.keys())[0]
        self._mode = mode
        self._terminal = is_probe_terminal
        self._efp = EFP(elect=elect, pol=polar, rep=repul, disp=disp, corr=correc, ccut=ccut, pcut=pcut, ecut=ecut, ea=True,
                        freq=True, cunit=True, mode=mode)
        self.mdinput_text = None; self.args = None; self.idx = None; self.near_atoms = None

    # ---------------------------------------------------------------- amide detection (biomol.py:331-438)
    def find_amide_atoms(self, top, xyz):
        """(far CONH groups, near-zone CONH groups, far CON- groups next to proline) as lists of atom-index lists; xyz in
        Angstrom.  A peptide group is four consecutive atoms named C, O, N, H (three without the H before proline) whose
        C-N distance is below 2.0 A; the near zone holds the groups with an atom within 2.4 A of the probe residue's CA."""
        names = top.name
        amides, amides_pro = [], []
        for i in range(len(names) - 2):
            if names[i] == 'C' and names[i + 1] == 'O' and names[i + 2] == 'N':
                if i + 3 < len(names) and names[i + 3] == 'H':
                    amides.append([i, i + 1, i + 2, i + 3])
                else:
                    amides_pro.append([i, i + 1, i + 2])
        dist = lambda a, b: float(numpy.linalg.norm(xyz[a] - xyz[b]))
        amides = [a for a in amides if dist(a[0], a[2]) <= 2.0]
        amides_pro = [a for a in amides_pro if dist(a[0], a[2]) <= 2.0]
        ca = [i for i in range(len(names)) if names[i] == 'CA' and top.resname[i] == self._probe_label]
        assert ca, ' No CA atom found in the probe residue %s' % self._probe_label
        d = numpy.linalg.norm(xyz[:, None, :] - xyz[ca][None, :, :], axis=-1).min(1)
        close_atoms = set(int(i) for i in numpy.nonzero(d < 2.4)[0]) - set(ca)
        near = [a for a in amides if any(x in close_atoms for x in a)]
        far = [a for a in amides if not any(x in close_atoms for x in a)]
        want = 1 if self._terminal else 2
        assert len(near) == want, (' It must be just %d CONH group(s) close to %s (found %d)' % (want, self._probe_label, len(near)))
        return far, near, amides_pro

    # ---------------------------------------------------------------- $Frag input (biomol.py:249-328)
    def create_solefp_inp(self, top, far, far_pro, conh='nma', include_proline_con=True):
        if conh not in _CONH_MODELS:
            raise NotImplementedError(' Other models of CONH group than <%s> are not implemented yet!' % conh)
        span, ro, sl, ro_pro, sl_pro = _CONH_MODELS[conh]
        residues = top.residues()
        log = ''; probe_key = None
        for rname, rid, atoms in residues:
            if rname == self._probe_label:
                probe_key = rname + str(rid)
                for frag in self._probe[self._probe_label]:
                    log += _frag_block(self._probe_label, frag, atoms + 1)
        assert probe_key is not None, ' Probe residue %s is not in the topology' % self._probe_label
        res_logs = {}
        nat = len(top)                               # (the reference writes `span` atoms from the carbonyl C whatever follows)
        for a in far:
            res_logs['CONH' + str(a)] = _frag_block('CONH', (conh, ro, sl), numpy.arange(a[0], min(a[0] + span, nat)) + 1)
        if include_proline_con:
            for a in far_pro:
                res_logs['CONX' + str(a)] = _frag_block('CONX', (conh, ro_pro, sl_pro), numpy.arange(a[0], min(a[0] + span, nat)) + 1)
        for rname, rid, atoms in residues:
            if rname == self._probe_label:
                continue
            if rname not in self._res:
                if self.lprint:
                    print(' WARNING: Residue %5s not found in topology conversion file' % rname)
                continue
            res_logs[rname + str(rid)] = ''.join(_frag_block(rname, frag, atoms + 1) for frag in self._res[rname])
        for key in sorted(res_logs):
            log += '! === %s RESIDUE === \n' % key + res_logs[key]
        return log

    def build(self, top, xyz, conh='nma', include_proline_con=True, out_inp=None):
        """One-off set-up from the topology and one structure (Angstrom): the $Frag input, its MDInput arguments, and the
        environment (near-zone CONH charges) appended as one-atom fragments.  Returns (args, idx)."""
        far, near, far_pro = self.find_amide_atoms(top, xyz)
        self.mdinput_text = self.create_solefp_inp(top, far, far_pro, conh, include_proline_con)
        if self.write_solefp_input and out_inp:
            with open(out_inp, 'w') as fh:
                fh.write(self.mdinput_text)
        (ind, nmol, bsm, supl, reord), idx = MDInput(self.mdinput_text).get()
        ind, nmol, bsm, supl, reord, idx = list(ind), list(nmol), list(bsm), list(supl), list(reord), list(idx)
        self.near_atoms = [int(x) for a in near for x in a]
        t0 = len(bsm)
        for q in CONH_CHARGES:                      # four environment types (C, O, N, H charges), one instance per near-zone atom
            bsm.append(point_charge_fragment(q, 'CONH point charge')); supl.append(None); reord.append(numpy.arange(1))
        for k, atom in enumerate(self.near_atoms):
            ind.append(t0 + k % 4); nmol.append(1); idx.append(atom)
        self.args = [ind, nmol, bsm, supl, reord]; self.idx = numpy.array(idx, int)
        self.n_far, self.n_near, self.n_far_pro = len(far), len(near), len(far_pro)
        return self.args, self.idx

    # ---------------------------------------------------------------- frame loop (biomol.py:150-231)
    def header(self):
        line = '%8s' % '#  Time'.ljust(8)
        line += ' %10s' * len(DW_LIST) % DW_LIST
        line += ' %10s' * len(DW_ENV_LIST) % DW_ENV_LIST
        return line + ' %10s %10s %10s\n' % ('RMS-C'.rjust(10), 'RMS-S'.rjust(10), 'RMS-AVE'.rjust(10))

    def run(self, top, traj, dframes=1, nframes=10, conh='nma', include_proline_con=True, out_inp='biomolecule.sol',
            times=None, batch=1024):
        """Analyse the MD trajectory: `top` a .gro / .pdb file or (Topology, xyz[Angstrom]); `traj` a trajectory file or an
        array [nframes, natoms, 3] (Angstrom; XTC files are nm).  Writes the report and returns (rows, status)."""
        from . import _lib
        if isinstance(top, str):
            top, xyz0 = Topology.read(top)
        else:
            top, xyz0 = top
        unit = AngstromToBohr
        if isinstance(traj, str) and traj.lower().endswith('.xtc'):
            unit = 10.0 * AngstromToBohr
        X = mdrun.load_trajectory(traj, len(top)) if isinstance(traj, str) else numpy.asarray(traj)
        if X.ndim == 2:
            X = X[None]
        X = X[:nframes]
        first = numpy.asarray(X[0], numpy.float64) * (unit / AngstromToBohr)
        self.build(top, first if first.shape == xyz0.shape else xyz0, conh, include_proline_con, out_inp)
        efp = self._efp
        rows_all, st_all = [], []
        with open(self.report, 'w') as out:
            out.write(self.header()); out.flush()
            for b0 in range(0, len(X), batch):
                blk = X[b0:b0 + batch]
                kw = dict(remove_clashes=True, rcl_algorithm=self.rcl_algorithm, include_ions=self.include_ions,
                          include_polar=self.include_polar)
                if blk.dtype == numpy.float32:
                    rows, status = efp.eval_frames(numpy.ascontiguousarray(blk), *self.args, scale=unit,
                                                   idx=self.idx.astype(numpy.int32), **kw)
                else:
                    rows, status = efp.eval_frames(numpy.ascontiguousarray(blk[:, self.idx, :] * unit), *self.args, **kw)
                dw = efp.shifts_of(rows)
                O = _lib.OUT
                for f in range(len(rows)):
                    t = float(times[b0 + f]) if times is not None else float(b0 + f)
                    log = '%8.4f' % t
                    log += len(DW_LIST) * ' %10.2f' % tuple(dw[k][f] for k in DW_LIST)
                    log += len(DW_ENV_LIST) * ' %10.2f' % tuple(dw[k][f] for k in DW_ENV_LIST)
                    log += ' %10.5f %10.5f %10.5f\n' % (rows[f, O['rms_c']], rows[f, O['rms_smax']], rows[f, O['rms_ave']])
                    out.write(log)
                out.flush()
                rows_all.append(numpy.array(rows)); st_all.append(numpy.array(status))
        return numpy.concatenate(rows_all), numpy.concatenate(st_all)
