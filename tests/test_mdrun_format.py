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


def _write_dcd(path, X, cell=False, endian='<'):
    import struct
    nf, na = X.shape[:2]
    with open(path, 'wb') as fh:
        ic = [0] * 20; ic[0] = nf; ic[1] = 1; ic[2] = 1; ic[10] = 1 if cell else 0; ic[19] = 24
        fh.write(struct.pack(endian + 'i', 84) + b'CORD' + struct.pack(endian + '20i', *ic) + struct.pack(endian + 'i', 84))
        title = b'REMARKS synthetic'.ljust(80)
        fh.write(struct.pack(endian + 'i', 84) + struct.pack(endian + 'i', 1) + title + struct.pack(endian + 'i', 84))
        fh.write(struct.pack(endian + 'i', 4) + struct.pack(endian + 'i', na) + struct.pack(endian + 'i', 4))
        for f in range(nf):
            if cell:
                fh.write(struct.pack(endian + 'i', 48) + struct.pack(endian + '6d', 30, 0, 30, 0, 0, 30) + struct.pack(endian + 'i', 48))
            for d in range(3):
                fh.write(struct.pack(endian + 'i', 4 * na) + X[f, :, d].astype(endian + 'f4').tobytes() + struct.pack(endian + 'i', 4 * na))


def test_dcd_and_mdcrd_readers(tmp_path):
    from slv_b200 import mdrun
    rng = np.random.default_rng(0)
    X = np.round(rng.normal(0, 10, (5, 7, 3)), 3)
    for cell in (False, True):
        for en in ('<', '>'):
            fn = str(tmp_path / ('t%d%s.dcd' % (cell, 'l' if en == '<' else 'b')))
            _write_dcd(fn, X, cell, en)
            Y = mdrun.read_dcd_frames(fn)
            assert Y.shape == X.shape and np.abs(Y - X).max() < 1e-5
            assert np.abs(mdrun.read_dcd_frames(fn, 2, 4) - X[2:4]).max() < 1e-5
    fn = tmp_path / 't.mdcrd'
    vals = X.reshape(5, -1)
    with open(fn, 'w') as fh:
        fh.write('title\n')
        for f in range(5):
            row = list(vals[f]) + [30.0, 30.0, 30.0]
            for i in range(0, len(row), 10):
                fh.write(''.join('%8.3f' % v for v in row[i:i + 10]) + '\n')
    Y = mdrun.read_mdcrd_frames(str(fn), 7, box=True)
    assert Y.shape == X.shape and np.abs(Y - X).max() < 1e-9
