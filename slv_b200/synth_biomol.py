"""Synthetic protein-like system for the biomolecular front-end (used by bench.py --config 4 and the tests; like
slv_b200/synth.py there is no network for real trajectories).

A chain of residues whose EFP fragments (side chains from a .tc table, peptide links as N-methylacetamide bodies, the
probe residue) are rigid parameter geometries in random orientations on a jittered lattice; the residue atom lists are
composed by inverting the .tc `reorder` maps, so every fragment superimposes onto its MD atoms exactly.  Backbone atoms
are named N, H, CA, HA ... C, O in GROMACS order, which is what the amide detection of biomol.py:331-438 keys on."""
import numpy as np

from . import synth
from .biomol import Topology
from .slvpar import Frag
from .units import AngstromToBohr

B2A = 1.0 / AngstromToBohr
_NMA_LINK = dict(C=2, O=4, N=1, H=7)          # 0-based parameter atoms of nma.frg taken by C, O (residue i) and N, H (residue i+1)


def _rot(rng):
    return synth.random_rotations(rng, 1)[0]


def make_system(tc, probe_tc, sequence, rng, nwater=0, spacing=6.0, frag_cache=None):
    """-> Topology, xyz [natoms,3] in Angstrom.  `sequence`: residue names (keys of tc); the probe residue is inserted in
    the middle.  Waters (residue SOL) fill further lattice cells."""
    fc = {} if frag_cache is None else frag_cache

    def frag(name):
        if name not in fc:
            fc[name] = Frag(name).get_pos() * B2A
        return fc[name]

    probe_label = list(probe_tc.keys())[0]
    seq = list(sequence[:len(sequence) // 2]) + [probe_label] + list(sequence[len(sequence) // 2:])
    pidx = len(sequence) // 2
    # lattice cells by distance from the origin; cell 0 region is kept for the probe and its two links
    # serpentine walk through a cubic grid: consecutive cells are neighbours, so the bodies of one residue (and of residues
    # adjacent in sequence) stay together as in a folded chain.  The walk is shifted so that the probe (at the origin) sits
    # where the chain passes when its turn comes; the cells around the origin are kept for the probe and its two links.
    def ncells(rn):
        return max(1, len(tc[rn]))
    nw_before = nwater // 2                               # half of the waters go on the walk before the chain, half after
    nbefore = nw_before + sum(ncells(r) for r in seq[:pidx]) + max(0, pidx - 1)
    nbody = sum(ncells(r) for r in seq if r != probe_label) + len(seq) + nwater + 64
    m = int(np.ceil(nbody ** (1.0 / 3.0))) + 3
    walk = []
    for ix in range(m):
        ys = range(m) if ix % 2 == 0 else range(m - 1, -1, -1)
        for ny, iy in enumerate(ys):
            zs = range(m) if (ix * m + ny) % 2 == 0 else range(m - 1, -1, -1)
            for iz in zs:
                walk.append((ix, iy, iz))
    lat = np.array(walk, float) * spacing
    start = (len(lat) - nbody) // 2                       # leave room on both ends of the walk
    lat = lat[start:] - lat[start + nbefore]
    lat = lat[np.linalg.norm(lat, axis=1) > 1.2 * spacing]
    cell = iter(lat)

    placed = {}                                           # lattice cell -> atoms already there (overlap checks)

    def key(c):
        return tuple(int(v) for v in np.round(np.asarray(c) / spacing))

    def place(pos, centre, tries=12):
        """The body in a random orientation around `centre`: of `tries` candidates the one that stays farthest from the
        atoms already placed in the neighbouring cells."""
        k0 = key(centre)
        near = [placed[(k0[0] + a, k0[1] + b, k0[2] + c)] for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)
                if (k0[0] + a, k0[1] + b, k0[2] + c) in placed]
        near = np.concatenate(near) if near else np.zeros((0, 3))
        best, best_d = None, -1.0
        for _ in range(tries if len(near) else 1):
            cand = (pos - pos.mean(0)) @ _rot(rng) + centre + rng.uniform(-0.3, 0.3, 3)
            d = np.linalg.norm(cand[:, None, :] - near[None, :, :], axis=-1).min() if len(near) else 1e9
            if d > best_d:
                best, best_d = cand, d
        placed[k0] = np.concatenate([placed[k0], best]) if k0 in placed else best
        return best

    # probe neighbourhood, hand-built: CA at the origin, the probe body above it, the two near-zone links beside it
    nma = frag('nma')
    link_prev = nma - nma[_NMA_LINK['N']] + np.array([-1.25, 0.75, 0.0])        # its N sits 1.46 A from CA
    link_next = nma - nma[_NMA_LINK['C']] + np.array([1.30, 0.78, 0.0])         # its C sits 1.52 A from CA
    pf = probe_tc[probe_label][0]
    pbody = frag(pf[0]); pbody = (pbody - pbody.mean(0)) + np.array([0.0, -0.6, 4.2])
    placed[(0, 0, 0)] = np.concatenate([link_prev, link_next, pbody])
    wat = frag('water')
    water_xyz = [place(wat, next(cell)) for _ in range(nw_before)]
    names, resid, resname, xyz = [], [], [], []
    links = {}
    for i, rn in enumerate(seq):
        frs = probe_tc[probe_label] if rn == probe_label else tc[rn]
        nat = max([6] + [int(max(ro)) + 2 for _, ro, _ in frs])
        centre = np.zeros(3) if rn == probe_label else next(cell)
        pos = centre + rng.normal(0, 0.5, (nat, 3))                 # filler for atoms no fragment uses
        for k, (fname, ro, sl) in enumerate(frs):
            body = pbody if rn == probe_label else place(frag(fname), centre if k == 0 else next(cell))
            for ip, kk in enumerate(ro):
                if kk > 0:
                    pos[kk - 1] = body[ip]
        if i < len(seq) - 1:                                        # the peptide link to the next residue
            links[i] = link_prev if i == pidx - 1 else (link_next if i == pidx else place(nma, next(cell)))
        an = ['N', 'H', 'CA', 'HA'] + ['X%d' % k for k in range(nat - 6)] + ['C', 'O']
        if rn == probe_label:
            pos[2] = 0.0                                            # CA
        if i > 0:                                                   # N, H from the link to the previous residue
            pos[0] = links[i - 1][_NMA_LINK['N']]; pos[1] = links[i - 1][_NMA_LINK['H']]
        if i < len(seq) - 1:                                        # C, O from the link to the next residue
            pos[nat - 2] = links[i][_NMA_LINK['C']]; pos[nat - 1] = links[i][_NMA_LINK['O']]
        names += an; resid += [i + 1] * nat; resname += [rn] * nat; xyz.append(pos)
    water_xyz += [place(wat, next(cell)) for _ in range(nwater - nw_before)]
    for w in range(nwater):
        xyz.append(water_xyz[w]); names += ['OW', 'HW1', 'HW2']; resid += [len(seq) + 1 + w] * 3; resname += ['SOL'] * 3
    return Topology(resid, resname, names), np.concatenate(xyz)


def jiggle(xyz, top, nframes, rng, amp=0.02):
    """A short 'trajectory': the structure with small independent atom displacements per frame (Angstrom)."""
    return xyz[None] + rng.normal(0, amp, (nframes,) + xyz.shape)
