"""Report writer of the MD frame loop (util/slv_md-run_nopp:127-153, 182-190): header and row formats."""
import numpy as np


def test_header_and_row_format():
    from slv_b200 import mdrun, _lib
    from slv_b200.solefp import shifts_from_rows
    assert mdrun.HEADER.split() == ['#', 'Frame', 'EL-CAMM', 'EL-MEA', 'EL-EA', 'CORR-MEA', 'CORR-EA', 'POL-MEA', 'POL-EA',
                                    'REP-MEA', 'REP-EA', 'DISP-MEA', 'DISP-EA', 'DISP-ISO', 'TOTAL', 'RMS-C', 'RMS-S', 'RMS-AVE']
    assert len(mdrun.HEADER) == 8 + 16 * 11
    rows = np.zeros((2, _lib.NOUT)); rows[:, :8] = np.arange(16).reshape(2, 8) * 1e-5
    rows[:, 8:11] = [[0.01, 0.02, 0.015], [0.011, 0.021, 0.016]]
    s = shifts_from_rows(rows, True)
    lines = mdrun.format_rows(rows, s, first_frame=7)
    assert lines[0].startswith('       7 ') and lines[1].startswith('       8 ')
    cols = lines[0].split()
    assert len(cols) == 17 and len(lines[0]) == 8 + 13 * 11 + 3 * 11 + 1
    assert abs(float(cols[1]) - round(s['solcamm'][0], 2)) < 1e-9 and abs(float(cols[13]) - round(s['total'][0], 2)) < 1e-9
    assert cols[14] == '0.01000' and cols[16] == '0.01500'


def test_read_xyz_frames(tmp_path):
    from slv_b200 import mdrun
    fn = tmp_path / 't.xyz'
    fn.write_text('2\nc\nO 0 0 0\nH 1 0 0\n2\nc\nO 0 0 1\nH 1 0 1\n')
    X = mdrun.read_xyz_frames(str(fn))
    assert X.shape == (2, 2, 3) and X[1, 1, 2] == 1.0
