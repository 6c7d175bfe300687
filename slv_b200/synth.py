"""Synthetic MD trajectories of the shapes BASELINE.json names.

Host-side NumPy generator used by bench.py and the tests (there is no network for real
trajectories).  One frame = solute (fragment 0) + N solvent fragments, coordinates in Bohr,
atoms of every instance contiguous and in parameter order (reorder lists are identity).

  solute : parameter geometry + N(0, jit_c) atom jitter, random rigid rotation per frame,
           centred at the origin
  solvent: rigid parameter geometry (+ N(0, jit_s) jitter), random orientation, centres on a
           jittered cubic lattice (spacing 5.87 Bohr ~ 1 g/cm^3 water); the N lattice points
           nearest to the solute are kept; any solvent with an atom closer than <rmin> Bohr to
           a solute atom is pushed radially outwards until it is not.
"""
import numpy


def random_rotations(rng, n):
    """n uniformly distributed proper rotation matrices [n,3,3] (from unit quaternions)."""
    q = rng.normal(size=(n, 4))
    q /= numpy.linalg.norm(q, axis=1)[:, None]
    w, x, y, z = q.T
    R = numpy.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def make_trajectory(solute_pos, solvent_pos_list, solvent_types, nframes, seed=1,
                    spacing=5.87, jit_c=0.02, jit_s=0.01, jit_lattice=0.6, rmin=3.0):
    """Returns coords [nframes, natoms_total, 3] (float64, Bohr).

    solute_pos        [nat_c,3] parameter geometry of the solute
    solvent_pos_list  list of parameter geometries, one per solvent TYPE
    solvent_types     int array [nsolv]: index into solvent_pos_list for every solvent instance
    """
    rng = numpy.random.default_rng(seed)
    solvent_types = numpy.asarray(solvent_types, int)
    nsolv = len(solvent_types)
    nat_c = len(solute_pos)
    nats = numpy.array([len(solvent_pos_list[t]) for t in solvent_types], int)
    offs = nat_c + numpy.concatenate([[0], numpy.cumsum(nats)])
    ntot = int(offs[-1])
    out = numpy.empty((nframes, ntot, 3))
    # lattice points sorted by distance from the origin (the solute sits there)
    m = int(numpy.ceil((3.0 * (nsolv + 30) / (4.0 * numpy.pi)) ** (1.0 / 3.0))) + 2
    g = numpy.arange(-m, m + 1)
    lat = numpy.stack(numpy.meshgrid(g, g, g, indexing='ij'), -1).reshape(-1, 3).astype(float) + 0.5
    lat = lat[numpy.argsort((lat * lat).sum(1), kind='stable')] * spacing
    sc = solute_pos - solute_pos.mean(0)
    cent = [p - p.mean(0) for p in solvent_pos_list]
    for f in range(nframes):
        Rc = random_rotations(rng, 1)[0]
        xc = sc @ Rc + rng.normal(0.0, jit_c, sc.shape)
        out[f, :nat_c] = xc
        centres = lat[:nsolv] + rng.uniform(-jit_lattice, jit_lattice, (nsolv, 3))
        Rs = random_rotations(rng, nsolv)
        for t in range(len(solvent_pos_list)):
            idx = numpy.nonzero(solvent_types == t)[0]
            if len(idx) == 0:
                continue
            x = numpy.einsum('ka,nab->nkb', cent[t], Rs[idx]) + centres[idx][:, None, :]
            x = x + rng.normal(0.0, jit_s, x.shape)
            # push away instances that clash with the solute
            for _ in range(40):
                d = numpy.linalg.norm(x[:, :, None, :] - xc[None, None, :, :], axis=-1).min((1, 2))
                bad = d < rmin
                if not bad.any():
                    break
                c = x[bad].mean(1)
                x[bad] += (0.4 * c / numpy.linalg.norm(c, axis=1)[:, None])[:, None, :]
            na = len(cent[t])
            for j, k in enumerate(idx):
                out[f, offs[k]:offs[k] + na] = x[j]
    return out
