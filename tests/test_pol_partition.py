"""CPU mirror of the integer logic of k_pol's cost-weighted sweep partition (slv_b200/csrc/k_pol.cuh, resident sweep):
the GPU parity tests exercise a handful of centre counts, this walks all of them."""
import pytest


def _partition(ncent, nwarp=8, rows=2):
    bsh = 6 if rows == 2 else 5
    nb2 = (ncent + (1 << bsh) - 1) >> bsh
    nfull = nb2 - 1; ufull = nfull * ncent
    clast = 1 if (rows == 2 and ncent - (nfull << bsh) <= 32) else 2
    F = 2 * ufull; W2 = F + clast * ncent
    Qc = (W2 + nwarp - 1) // nwarp; Qc += Qc & 1

    def unit_of(x):
        x = min(x, W2)
        return x >> 1 if x <= F else ufull + (x - F) // clast

    def warp_of(u):
        return (2 * u if u < ufull else F + (u - ufull) * clast) // Qc
    return nb2, unit_of, warp_of, Qc


@pytest.mark.parametrize('rows', [1, 2])
def test_sweep_ranges_tile_the_pair_space_and_the_combination_reads_what_was_written(rows):
    nwarp = 8
    for ncent in range(1, 1025):
        nb2, unit_of, warp_of, Qc = _partition(ncent, nwarp, rows)
        W = nb2 * ncent
        bounds = [unit_of(w * Qc) for w in range(nwarp + 1)]
        assert bounds[0] == 0 and bounds[-1] == W and all(a <= b for a, b in zip(bounds, bounds[1:])), ncent
        split = any(unit_of(w * Qc) % ncent for w in range(1, nwarp))
        finished, written = {}, {}                      # block -> warp that finished it in the sweep; (warp, slot) -> block
        for w in range(nwarp):
            u, u1 = bounds[w], bounds[w + 1]
            while u < u1:
                b = u // ncent; ja = u - b * ncent; jb = min(ncent, ja + (u1 - u))
                if ja == 0 and jb == ncent:
                    assert b not in finished
                    finished[b] = w
                else:
                    slot = 0 if ja > 0 else 1
                    assert (w, slot) not in written, (ncent, w, slot)      # one head and one tail partial per warp at most
                    written[(w, slot)] = b
                u += jb - ja
        if not split:
            assert not written and len(finished) == nb2, ncent
        for b in range(nb2):                              # the combination phase of the kernel
            wf, wl = warp_of(b * ncent), warp_of(b * ncent + ncent - 1)
            if wf == wl:
                assert finished.get(b) == wf, (ncent, b)
                continue
            assert b not in finished, (ncent, b)
            for w2 in range(wf, wl + 1):
                slot = 0 if b * ncent < unit_of(w2 * Qc) else 1
                assert written.get((w2, slot)) == b, (ncent, b, w2, slot)
        assert sorted(list(finished) + sorted(set(written.values()))) == list(range(nb2)), ncent
