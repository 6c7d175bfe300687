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


def make_system(tc, probe_tc, sequence, rng, nwater=0, spacing=9.0, frag_cache=None):
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
    m = int(np.ceil((2 * len(seq) + nwater + 40) ** (1.0 / 3.0))) + 2
    g = np.arange(-m, m + 1)
    lat = np.stack(np.meshgrid(g, g, g, indexing='ij'), -1).reshape(-1, 3).astype(float) * spacing
    lat = lat[np.argsort((lat * lat).sum(1), kind='stable')]
    lat = lat[np.linalg.norm(lat, axis=1) > 1.2 * spacing]
    cell = iter(lat)

    def place(pos, centre):
        return (pos - pos.mean(0)) @ _rot(rng) + centre + rng.uniform(-0.4, 0.4, 3)

    # probe neighbourhood, hand-built: CA at the origin, the probe body above it, the two near-zone links beside it
    nma = frag('nma')
    bodies = {}
    link_prev = nma - nma[_NMA_LINK['N']] + np.array([-1.25, 0.75, 0.0])        # its N sits 1.46 A from CA
    link_next = nma - nma[_NMA_LINK['C']] + np.array([1.30, 0.78, 0.0])         # its C sits 1.52 A from CA
    pf = probe_tc[probe_label][0]
    pbody = frag(pf[0]); pbody = (pbody - pbody.mean(0)) + np.array([0.0, -0.6, 4.2])
    names, resid, resname, xyz = [], [], [], []
    links = {}
    for i in range(len(seq) - 1):
        if i == pidx - 1:
            links[i] = link_prev
        elif i == pidx:
            links[i] = link_next
        else:
            links[i] = place(nma, next(cell))
    for i, rn in enumerate(seq):
        frs = probe_tc[probe_label] if rn == probe_label else tc[rn]
        nat = max([6] + [int(max(ro)) + 2 for _, ro, _ in frs])
        centre = np.zeros(3) if rn == probe_label else next(cell)
        pos = centre + rng.normal(0, 0.5, (nat, 3))                 # filler for atoms no fragment uses
        for fname, ro, sl in frs:
            body = pbody if rn == probe_label else place(frag(fname), centre if len(frs) == 1 else next(cell))
            for ip, k in enumerate(ro):
                if k > 0:
                    pos[k - 1] = body[ip]
        an = ['N', 'H', 'CA', 'HA'] + ['X%d' % k for k in range(nat - 6)] + ['C', 'O']
        if rn == probe_label:
            pos[2] = 0.0                                            # CA
        if i > 0:                                                   # N, H from the link to the previous residue
            pos[0] = links[i - 1][_NMA_LINK['N']]; pos[1] = links[i - 1][_NMA_LINK['H']]
        if i < len(seq) - 1:                                        # C, O from the link to the next residue
            pos[nat - 2] = links[i][_NMA_LINK['C']]; pos[nat - 1] = links[i][_NMA_LINK['O']]
        names += an; resid += [i + 1] * nat; resname += [rn] * nat; xyz.append(pos)
    wat = frag('water')
    for w in range(nwater):
        xyz.append(place(wat, next(cell))); names += ['OW', 'HW1', 'HW2']; resid += [len(seq) + 1 + w] * 3; resname += ['SOL'] * 3
    return Topology(resid, resname, names), np.concatenate(xyz)


def jiggle(xyz, top, nframes, rng, amp=0.02):
    """A short 'trajectory': the structure with small independent atom displacements per frame (Angstrom)."""
    return xyz[None] + rng.normal(0, amp, (nframes,) + xyz.shape)
