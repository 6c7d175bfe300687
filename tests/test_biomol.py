"""Protein fragmentation front-end (solvshift/biomol.py): .tc parser, $Frag generation, amide zones (CPU), and the batched
evaluation with near-zone point charges + clash removal against the oracle (GPU)."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TC = os.path.join(ROOT, 'slv_b200', 'data', 'tc')


def _system(seed=4, nwater=20, seq=('ASP', 'SER', 'GLY', 'LYS', 'ASN', 'ALA', 'GLU', 'TYR', 'PHE', 'THR')):
    from slv_b200.synth_biomol import make_system
    from slv_b200 import biomol
    tc = biomol.parse_tc(os.path.join(TC, 'gmx.tc')); ptc = biomol.parse_tc(os.path.join(TC, 'probe.tc'))
    top, xyz = make_system(tc, ptc, list(seq), np.random.default_rng(seed), nwater=nwater)
    return tc, ptc, top, xyz


def test_tc_parser_and_frag_input(tmp_path):
    from slv_b200 import biomol
    tc, ptc, top, xyz = _system()
    assert set(('ASP', 'SOL', 'GLY', 'LYS', 'NA+')) <= set(tc) and tc['GLY'] == [] and list(ptc) == ['SCN']
    name, ro, sl = tc['ASP'][0]
    assert name == 'mecoo-' and list(ro) == [5, 6, 8, 7, 0, 9, 10] and list(sl) == [3, 6, 7]
    assert [f[0] for f in tc['ASN']] == ['chonh2', 'methane']          # commented-out alternatives are skipped
    b = biomol.BiomoleculeFragmentation(tc, ptc, 4, report=str(tmp_path / 'r.dat'), repul=False)
    far, near, pro = b.find_amide_atoms(top, xyz)
    assert len(near) == 2 and len(far) == 8 and pro == []
    for a in far + near:
        assert [top.name[i] for i in a] == ['C', 'O', 'N', 'H']
    (ind, nmol, bsm, supl, reord), idx = b.build(top, xyz)
    assert len(ind) == len(nmol) and sum(nmol) == len(idx) and ind[0] == 0 and bsm[0].get()['name'] == 'Methyl Thiocyanate'
    # identical blocks share one type; the last eight instances are the near-zone point charges on their frame atoms
    assert len(bsm) < len(ind) and [bsm[t].get().get('env') for t in ind[-8:]] == [True] * 8
    assert list(idx[-8:]) == [x for a in near for x in a]
    q = [float(bsm[t].get()['dmac'][0]) for t in ind[-8:]]
    assert q == list(biomol.CONH_CHARGES) * 2
    assert b.mdinput_text.startswith('$Frag\nSCN') and '! === ASP1 RESIDUE === ' in b.mdinput_text
    # topology readers
    gro = tmp_path / 's.gro'
    with open(str(gro), 'w') as fh:
        fh.write('t\n%5d\n' % len(top))
        for i in range(len(top)):
            fh.write('%5d%-5s%5s%5d%8.3f%8.3f%8.3f\n' % (top.resid[i], top.resname[i], top.name[i], (i + 1) % 100000, *(xyz[i] / 10.0)))
        fh.write('   1.0   1.0   1.0\n')
    t2, x2 = biomol.Topology.read(str(gro))
    assert list(t2.name) == list(top.name) and list(t2.resname) == list(top.resname) and np.abs(x2 - xyz).max() < 6e-3


@pytest.mark.gpu
def test_biomol_run_matches_oracle(tmp_path):
    """BASELINE config 4, small: the whole front-end -> one batched call (SolEFP with clash removal + near-zone CONH
    charges as environment fragments) against the oracle's eval_frame + eval_dma_frame on the same frames."""
    from slv_b200.synth_biomol import jiggle
    from slv_b200 import biomol
    from slv_b200.units import AngstromToBohr
    from oracle import solefp_oracle as O
    tc, ptc, top, xyz = _system()
    X = jiggle(xyz, top, 3, np.random.default_rng(9))
    rep = str(tmp_path / 'report.dat')
    b = biomol.BiomoleculeFragmentation(tc, ptc, 4, ccut=45.0, pcut=16.0, ecut=13.0, report=rep, include_ions=False, include_polar=False)
    rows, status = b.run((top, xyz), X, nframes=3)
    assert int(status.max()) == 0
    lines = open(rep).read().split('\n')
    assert lines[0].split()[:3] == ['#', 'Time', 'solcamm'] and len(lines[0].split()) == 2 + 13 + 6 + 3 and len(lines[1].split()) == 23
    ind, nmol, bsm, supl, reord = b.args
    dw = b._efp.shifts_of(rows)
    c = O.HartreePerHbarToCmRec
    nenv = 8
    pars = [x.get() for x in bsm]
    rem = b._efp._removable
    assert rem and all(pars[t]['name'] in ('Methane', 'Formamide', 'Acetate anion (-1)', 'N-methylacethamide', 'Methyl amonium cation (+1)',
                                           'Benzene', 'Methanol', 'Phenol') for t in rem)
    for f in range(3):
        fr = X[f][b.idx] * AngstromToBohr
        n_efp = sum(nmol[:-nenv])
        ref = O.eval_frame(fr[:n_efp], ind[:-nenv], nmol[:-nenv], pars, supl, reord, 4, 45.0, 16.0, 13.0, elect=True, ea=True, corr=True,
                           pol=True, disp=True, rep=True, removable=rem)
        env = np.concatenate([fr[n_efp:], np.array(biomol.CONH_CHARGES * 2)[:, None]], axis=1)
        pos_c = O.reorder_xyz(fr[:nmol[0]], reord[0])
        renv = O.eval_dma_frame(pos_c, pars[0], supl[0], 4, env)
        for k in ('ele_mea', 'ele_ea', 'cor_mea', 'cor_ea', 'pol_mea', 'dis_mea', 'dis_mea_iso', 'rep_mea'):
            assert abs(dw[k][f] - ref[k] * c) < 1e-6 * max(1.0, abs(ref[k] * c) * 1e-3), (f, k, dw[k][f], ref[k] * c)
        for k in ('ele_env_mea', 'ele_env_ea', 'cor_env_mea', 'cor_env_ea'):
            assert abs(dw[k][f] - renv[k] * c) < 1e-6 * max(1.0, abs(renv[k] * c) * 1e-3), (f, k, dw[k][f], renv[k] * c)
        tot_env = sum(renv[k] for k in renv) * c
        assert abs(dw['total+env'][f] - (dw['total'][f] + tot_env)) < 1e-6
        assert abs(float(lines[1 + f].split()[18]) - round(tot_env, 2)) < 0.011
