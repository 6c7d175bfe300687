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


def test_xtc_reader_roundtrip(tmp_path):
    """The library's XTC decoder (slv_b200/csrc/xtc_reader.h, host code: no GPU needed) against files written by the
    test-side transcription of the published xdr3dcoord compressor (tests/xtc_writer.py): water-like molecules (runs of
    small triples with the swapped first pair), a dilute gas (large triples only), coordinate ranges beyond 2^24 grid
    steps (the per-axis bit-size branch), nine atoms or fewer (uncompressed floats), frame sub-ranges."""
    from slv_b200 import mdrun, _lib
    from xtc_writer import write_xtc
    rng = np.random.default_rng(8)
    # 1. liquid-like: 300 three-atom molecules, atoms within 0.1 nm of each other
    cent = rng.uniform(0, 4, (7, 300, 1, 3))
    liquid = (cent + rng.normal(0, 0.04, (7, 300, 3, 3))).reshape(7, 900, 3)
    # 2. gas: no two consecutive atoms close
    gas = rng.uniform(-20, 20, (3, 50, 3))
    # 3. huge box: per-axis bit sizes (range > 0xffffff grid steps at precision 1000)
    huge = rng.uniform(-9000, 9000, (2, 40, 3)); huge[:, 1::2] = huge[:, ::2] + rng.normal(0, 0.02, (2, 20, 3))
    # 4. tiny system
    tiny = rng.uniform(0, 1, (4, 9, 3))
    for name, X, prec in (('liquid', liquid, 1000.0), ('liquid100', liquid, 100.0), ('gas', gas, 1000.0), ('huge', huge, 1000.0),
                          ('tiny', tiny, 1000.0)):
        fn = str(tmp_path / (name + '.xtc'))
        write_xtc(fn, X, precision=prec)
        Y = mdrun.read_xtc_frames(fn)
        assert Y.dtype == np.float32 and Y.shape == X.shape, name
        if name == 'tiny':
            assert np.array_equal(Y, X.astype(np.float32))
        else:
            q = np.where(X >= 0, np.floor(X * prec + 0.5), np.ceil(X * prec - 0.5))
            assert np.array_equal(Y, (q.astype(np.int64).astype(np.float32) * np.float32(1.0 / np.float32(prec)))), name
        assert np.array_equal(mdrun.read_xtc_frames(fn, 1, 2), Y[1:2])
        assert np.array_equal(mdrun.load_trajectory(fn), Y)
    import pytest
    bad = tmp_path / 'bad.xtc'
    bad.write_bytes(b'\x00\x00\x00\x01' * 20)
    with pytest.raises(_lib.SlvGpuError):
        mdrun.read_xtc_frames(str(bad))
    trunc = tmp_path / 'trunc.xtc'
    trunc.write_bytes(open(str(tmp_path / 'liquid.xtc'), 'rb').read()[:3000])
    with pytest.raises(_lib.SlvGpuError):
        mdrun.read_xtc_frames(str(trunc))
