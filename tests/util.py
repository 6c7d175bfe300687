"""Shared helpers of the test-suite (not product code)."""
import numpy as np

BohrToAngstrom = 0.529177249


def scan_geometries(golden, npoints=21):
    """The 21 geometries of examples/1_small-clusters/scan.py (Bohr): list of (15,3) arrays."""
    xn = np.array([r[1:] for r in golden['nma_xyz']]) / BohrToAngstrom
    xw = np.array([r[1:] for r in golden['water_xyz']]) / BohrToAngstrom
    xn = xn - xn.mean(0); xw = xw - xw.mean(0)
    r = xn[7] - xn[1]
    xw = xw + 6.0 * r / np.linalg.norm(r)
    t = 0.1
    rm = np.array([np.cos(t), -np.sin(t), 0, np.sin(t), np.cos(t), 0, 0, 0, 1]).reshape(3, 3)
    out = []
    for i in range(npoints):
        out.append(np.concatenate([xn, xw]))
        xw = xw @ rm
    return out


SCAN_SUPL = [np.array([2, 8, 3, 5]) - 1, np.array([1, 2, 3]) - 1]
SCAN_REORD = [np.arange(12), np.arange(3)]
